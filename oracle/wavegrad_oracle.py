"""CPU oracle for the cfg-4 hot path: WaveGrad denoiser + the spectrogram-conditioned sampling loop.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file (see oracle/sddm_oracle.py for the rule).

Functional restatement (plain fp32 torch ops on CPU, driven by a reference-layout ``state_dict``) of

* PositionalEncoding ........ /root/reference/model/wavegrad.py:20-49   (no 5000x scale; added to every time step)
* FiLM ...................... /root/reference/model/wavegrad.py:52-71
* UBlock .................... /root/reference/model/wavegrad.py:74-112
* DBlock .................... /root/reference/model/wavegrad.py:115-137
* WaveGrad.forward .......... /root/reference/model/wavegrad.py:140-179 (audio [B,T], noise_scale [B] / [B,1,1])
* SDDM_spectrogram.infer .... /root/reference/model/model.py:206-257.  As shipped the wrapper passes x_t [B,1,T] straight to
  WaveGrad.forward, whose ``audio.unsqueeze(1)`` then feeds a 4-D tensor to Conv1d and raises; the loop here
  is that loop with the one squeeze / unsqueeze that makes it run, and the golden generator applies the same adapter to the
  real reference modules.

Parity status: PINNED by tests/golden/make_golden_wavegrad.py -> tests/golden/wavegrad.npz (tests/test_wavegrad.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

DOWN = [(32, 128, 2), (128, 128, 2), (128, 256, 3), (256, 512, 5)]                      # DBlock(input, hidden, factor)
FILM = [(32, 128), (128, 128), (128, 256), (256, 512), (512, 512)]                      # FiLM(input, output)
UP = [(768, 512, 5, (1, 2, 1, 2)), (512, 512, 5, (1, 2, 1, 2)), (512, 256, 3, (1, 2, 4, 8)),
      (256, 128, 2, (1, 2, 4, 8)), (128, 128, 2, (1, 2, 4, 8))]                         # UBlock(input, hidden, factor, dilation)
HOP = 2 * 2 * 3 * 5 * 5                                                                  # 300 samples per spectrogram frame


def positional_encoding(dim: int, noise_level: Tensor) -> Tensor:
    """[N] -> [N, dim]   (wavegrad.py:44-49)."""
    count = dim // 2
    step = torch.arange(count, dtype=noise_level.dtype) / count
    enc = noise_level.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
    return torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)


def conv(sd, key, x, dilation=1, padding=0):
    return F.conv1d(x, sd[key + ".weight"], sd[key + ".bias"], dilation=dilation, padding=padding)


def film(sd: Dict[str, Tensor], i: int, x: Tensor, noise_level: Tensor):
    p = f"film.{i}."
    x = F.leaky_relu(conv(sd, p + "input_conv", x, padding=1), 0.2)
    x = x + positional_encoding(x.shape[1], noise_level)[:, :, None]
    return torch.chunk(conv(sd, p + "output_conv", x, padding=1), 2, dim=1)


def dblock(sd: Dict[str, Tensor], i: int, x: Tensor, factor: int) -> Tensor:
    p = f"downsample.{i}."
    size = x.shape[-1] // factor
    residual = F.interpolate(conv(sd, p + "residual_dense", x), size=size)
    x = F.interpolate(x, size=size)
    for j, d in enumerate((1, 2, 4)):
        x = conv(sd, p + f"conv.{j}", F.leaky_relu(x, 0.2), dilation=d, padding=d)
    return x + residual


def ublock(sd: Dict[str, Tensor], i: int, x: Tensor, shift: Tensor, scale: Tensor, factor: int, dil) -> Tensor:
    p = f"upsample.{i}."
    size = x.shape[-1] * factor
    b1 = conv(sd, p + "block1", F.interpolate(x, size=size))
    b2 = F.interpolate(F.leaky_relu(x, 0.2), size=size)
    b2 = conv(sd, p + "block2.0", b2, dilation=dil[0], padding=dil[0])
    b2 = F.leaky_relu(shift + scale * b2, 0.2)
    b2 = conv(sd, p + "block2.1", b2, dilation=dil[1], padding=dil[1])
    x = b1 + b2
    b3 = F.leaky_relu(shift + scale * x, 0.2)
    b3 = conv(sd, p + "block3.0", b3, dilation=dil[2], padding=dil[2])
    b3 = F.leaky_relu(shift + scale * b3, 0.2)
    b3 = conv(sd, p + "block3.1", b3, dilation=dil[3], padding=dil[3])
    return x + b3


def wavegrad_forward(sd: Dict[str, Tensor], spectrogram: Tensor, audio: Tensor, noise_scale: Tensor, trace: Optional[dict] = None) -> Tensor:
    """spectrogram [B,128,F], audio [B,300 F], noise_scale [B] (or [B,1,1]) -> eps_hat [B,300 F]  (wavegrad.py:167-179;
    the reference's final torch.squeeze also drops the batch dimension when B == 1 — callers reshape)."""
    nl = noise_scale.reshape(-1)
    x = audio.unsqueeze(1)
    films = []
    for i in range(5):
        x = conv(sd, "downsample.0", x, padding=2) if i == 0 else dblock(sd, i, x, DOWN[i - 1][2])
        films.append(film(sd, i, x, nl))
        if trace is not None:
            trace[f"d{i}"] = x
    x = conv(sd, "first_conv", spectrogram, padding=1)
    for i, (shift, scale) in enumerate(reversed(films)):
        x = ublock(sd, i, x, shift, scale, UP[i][2], UP[i][3])
        if trace is not None:
            trace[f"u{i}"] = x
    return conv(sd, "last_conv", x, padding=1).squeeze(1)


def sample_spectrogram(sd: Dict[str, Tensor], sched: Dict[str, Tensor], spectrogram: Tensor, noises: Tensor,
                       noise_condition: str = "sqrt_alpha_bar", eps_trace: Optional[List[Tensor]] = None) -> Tensor:
    """SDDM_spectrogram.infer (model.py:212-257, non-continuous) around WaveGrad with injected noise; returns [B,1,T]."""
    T = sched["betas"].numel() - 1
    B = spectrogram.shape[0]
    x_t = noises[0].reshape(B, 1, HOP * spectrogram.shape[-1]).clone()
    for t in range(T, 0, -1):
        level = (sched["sqrt_alpha_bar"][t] if noise_condition == "sqrt_alpha_bar" else torch.tensor(float(t))) * torch.ones(B)
        eps = wavegrad_forward(sd, spectrogram, x_t.squeeze(1), level).reshape(x_t.shape)
        if eps_trace is not None:
            eps_trace.append(eps)
        x = (x_t - sched["predicted_noise_coeff"][t] * eps) / sched["alphas"][t] ** 0.5
        if t > 1:
            x = x + sched["sigma"][t] * noises[T + 1 - t].reshape(x.shape)
        x_t = x.clamp_(-1.0, 1.0)
    return x_t
