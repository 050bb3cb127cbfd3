"""CPU oracle for the cfg-5 hot path: DiffWave denoiser + the spectrogram-conditioned sampling loop.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of the bench tools may import it, and only as the
checker / the CPU baseline — never as the thing shipped.

Functional restatement (plain fp32 torch ops on CPU, driven by a reference-layout ``state_dict``) of

* DiffusionEmbedding ........ /root/reference/model/diffwave.py:22-45   (note :28: vector 10^((k/64)*4/63), k < 64)
* SpectrogramUpsampler ...... /root/reference/model/diffwave.py:48-61   (2 x ConvTranspose2d [3,32] / stride [1,16] / pad [1,8])
* ResidualBlock ............. /root/reference/model/diffwave.py:64-108  (split=True branch, the constructor default)
* DiffWave.forward .......... /root/reference/model/diffwave.py:133-155
* SDDM_spectrogram.infer .... /root/reference/model/model.py:206-257    (pure-noise start, p_transition 'original')
* p_transition .............. /root/reference/model/diffusion.py:177-190

Parity status: PINNED.  ``tests/golden/make_golden_diffwave.py`` imports the real reference modules from
/root/reference (build container only) and commits golden vectors (``tests/golden/diffwave.npz``);
``tests/test_diffwave.py`` checks this restatement against them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

DIFFWAVE_DEFAULT_CFG = dict(freq_bins=513, residual_channels=64, residual_layers=30, dilation_cycle_length=10)


def embedding_vector(dim: int = 128) -> Tensor:
    """diffwave.py:26-28 (the commented-out exp(-log(1e4) step) form is NOT what runs)."""
    step = torch.arange(dim // 2) / (dim // 2)
    return 10.0 ** (step * 4.0 / 63)


def diffusion_embedding(sd: Dict[str, Tensor], diffusion_step: Tensor) -> Tensor:
    """[B,1] step values -> [B,512]   (diffwave.py:33-45)."""
    enc = diffusion_step * embedding_vector()
    enc = torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)
    x = F.linear(enc, sd["diffusion_embedding.projection1.weight"], sd["diffusion_embedding.projection1.bias"])
    x = x * torch.sigmoid(x)
    x = F.linear(x, sd["diffusion_embedding.projection2.weight"], sd["diffusion_embedding.projection2.bias"])
    return x * torch.sigmoid(x)


def spectrogram_upsampler(sd: Dict[str, Tensor], spec: Tensor) -> Tensor:
    """[B,F,frames] -> [B,F,256*frames]   (diffwave.py:54-61)."""
    x = spec.unsqueeze(1)
    for i in (1, 2):
        x = F.conv_transpose2d(x, sd[f"spectrogram_upsampler.conv{i}.weight"], sd[f"spectrogram_upsampler.conv{i}.bias"],
                               stride=(1, 16), padding=(1, 8))
        x = F.leaky_relu(x, 0.4)
    return x.squeeze(1)


def residual_block(sd: Dict[str, Tensor], i: int, x: Tensor, cond_up: Tensor, emb: Tensor, dilation: int) -> Tuple[Tensor, Tensor]:
    """diffwave.py:85-108, split=True."""
    p = f"residual_layers.{i}."
    e = F.linear(emb, sd[p + "diffusion_projection.weight"], sd[p + "diffusion_projection.bias"]).unsqueeze(-1)
    c = F.conv1d(cond_up, sd[p + "conditioner_projection.weight"], sd[p + "conditioner_projection.bias"])
    y = F.conv1d(x + e, sd[p + "dilated_conv.weight"], sd[p + "dilated_conv.bias"], padding=dilation, dilation=dilation) + c
    gate, filt = torch.chunk(y, 2, dim=1)
    y = torch.sigmoid(gate) * torch.tanh(filt)
    residual = F.conv1d(y, sd[p + "output_residual.weight"], sd[p + "output_residual.bias"])
    skip = F.conv1d(y, sd[p + "output_projection.weight"], sd[p + "output_projection.bias"])
    return (x + residual) / math.sqrt(2.0), skip


def diffwave_forward(sd: Dict[str, Tensor], spectrogram: Tensor, audio: Tensor, diffusion_step: Tensor,
                     residual_layers: int = 30, dilation_cycle_length: int = 10, trace: Optional[dict] = None) -> Tensor:
    """spectrogram [B,F,frames], audio [B,1,T], diffusion_step [B,1,1] -> eps_hat [B,1,T]   (diffwave.py:133-155)."""
    step = diffusion_step.squeeze(-1)
    x = F.relu(F.conv1d(audio, sd["input_projection.weight"], sd["input_projection.bias"]))
    emb = diffusion_embedding(sd, step)
    up = spectrogram_upsampler(sd, spectrogram)
    if trace is not None:
        trace["upsampled"] = up
        trace["embedding"] = emb
    skip_sum = None
    for i in range(residual_layers):
        x, s = residual_block(sd, i, x, up, emb, 2 ** (i % dilation_cycle_length))
        skip_sum = s if skip_sum is None else skip_sum + s
        if trace is not None:
            trace[f"x{i}"] = x
    x = skip_sum / math.sqrt(residual_layers)
    if trace is not None:
        trace["skip"] = x
    x = F.relu(F.conv1d(x, sd["skip_projection.weight"], sd["skip_projection.bias"]))
    return F.conv1d(x, sd["output_projection.weight"], sd["output_projection.bias"])


def sample_spectrogram(sd: Dict[str, Tensor], sched: Dict[str, Tensor], spectrogram: Tensor, noises: Tensor, hop_samples: int,
                       noise_condition: str = "time_step", residual_layers: int = 30, dilation_cycle_length: int = 10,
                       eps_trace: Optional[List[Tensor]] = None) -> Tensor:
    """SDDM_spectrogram.infer (model.py:212-257, non-continuous) with injected noise:
    noises[0] -> x_T (the torch.randn of :216), noises[k] -> step t = T + 1 - k (randn_like of diffusion.py:187)."""
    T = sched["betas"].numel() - 1
    B = spectrogram.shape[0]
    x_t = noises[0].reshape(B, 1, hop_samples * spectrogram.shape[-1]).clone()
    for t in range(T, 0, -1):
        if noise_condition == "sqrt_alpha_bar":
            level = sched["sqrt_alpha_bar"][t] * torch.ones(B, 1, 1)
        elif noise_condition == "time_step":
            level = t * torch.ones(B, 1, 1)
        else:
            raise NotImplementedError
        eps = diffwave_forward(sd, spectrogram, x_t, level, residual_layers, dilation_cycle_length)
        if eps_trace is not None:
            eps_trace.append(eps)
        x = (x_t - sched["predicted_noise_coeff"][t] * eps) / sched["alphas"][t] ** 0.5      # diffusion.py:184
        if t > 1:
            x = x + sched["sigma"][t] * noises[T + 1 - t].reshape(x.shape)                    # diffusion.py:186-188
        x_t = x.clamp_(-1.0, 1.0)                                                            # diffusion.py:190
    return x_t
