"""CPU oracle (TEST INFRASTRUCTURE ONLY) of the STFT feature front-end: a plain restatement of what the reference computes in
prepare_spectrogram.py:20-55 through torchaudio 2.11 ``functional.spectrogram`` -> ``torch.stft`` and ``MelScale``.

Third-party arithmetic: torchaudio / torch.stft are not part of /root/reference; their published algorithm is restated here
(reflect padding of n_fft/2, frames of n_fft at stride hop, periodic window, onesided DFT, division by sqrt(sum w^2), magnitude,
HTK triangular filterbank) and pinned against outputs of torchaudio itself (tests/golden/stft.npz, written by
tests/golden/make_golden_stft.py in the build container)."""
import math

import numpy as np
import torch


def window(kind: str, n_fft: int) -> torch.Tensor:
    n = torch.arange(n_fft, dtype=torch.float64)
    if kind == "hamming":                                          # torch.hamming_window(periodic=True)
        return (0.54 - 0.46 * torch.cos(2.0 * math.pi * n / n_fft)).float()
    if kind == "hann":                                             # torch.hann_window(periodic=True)
        return (0.5 - 0.5 * torch.cos(2.0 * math.pi * n / n_fft)).float()
    raise NotImplementedError(kind)


def spectrogram(wav: torch.Tensor, n_fft: int, hop: int, kind: str) -> torch.Tensor:
    """[B, L] -> [B, n_fft/2+1, 1 + L//hop]: |STFT| / sqrt(sum w^2)   (torchaudio functional.spectrogram, power=1, normalized=True)."""
    w = window(kind, n_fft).double()
    x = wav.double()
    B, L = x.shape
    pad = n_fft // 2
    idx = torch.arange(-pad, L + pad)
    idx = torch.where(idx < 0, -idx, idx)
    idx = torch.where(idx >= L, 2 * (L - 1) - idx, idx)           # reflect padding (torch.stft center=True)
    xp = x[:, idx]
    frames = 1 + L // hop
    starts = torch.arange(frames) * hop
    seg = xp[:, starts[:, None] + torch.arange(n_fft)[None, :]] * w          # [B, frames, n_fft]
    X = torch.fft.rfft(seg, dim=-1)                                           # [B, frames, n_fft/2+1]
    S = X.abs() / torch.sqrt((w * w).sum())
    return S.transpose(1, 2).float()


def melscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks, norm=None, mel_scale='htk'."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def mel_spectrogram(wav, n_fft, hop, n_mels, sample_rate, f_min=20.0, f_max=None):
    S = spectrogram(wav, n_fft, hop, "hann")
    fb = melscale_fbanks(n_fft // 2 + 1, f_min, f_max if f_max is not None else sample_rate / 2.0, n_mels, sample_rate)
    return torch.matmul(S.transpose(-1, -2), fb).transpose(-1, -2)


def log_clamp(x: torch.Tensor) -> torch.Tensor:                   # prepare_spectrogram.py:41-44,50-53
    return torch.clamp((torch.log10(x) - 1 + 5) / 5, 0.0, 1.0)


def istft(spec: torch.Tensor, n_fft: int, hop: int, win: torch.Tensor, length: int) -> torch.Tensor:
    """Restatement of torch.istft(spec, n_fft, hop, window=win, center=True, normalized=False, onesided=True, length=length) for a complex
    [B, n_fft/2+1, frames] spectrogram: irfft per frame, window, overlap-add in ascending frame order, division by the window
    envelope, centre trim.  (torch's arithmetic: aten/src/ATen/native/SpectralOps.cpp istft; pinned by tests/golden/stft.npz 'istft.*')"""
    B, _, frames = spec.shape
    fr = torch.fft.irfft(spec, n=n_fft, dim=1) * win.reshape(1, -1, 1)                  # [B, n_fft, frames]
    total = n_fft + hop * (frames - 1)
    y = torch.zeros(B, total, dtype=fr.dtype)
    env = torch.zeros(total, dtype=fr.dtype)
    for m in range(frames):
        y[:, m * hop:m * hop + n_fft] += fr[:, :, m]
        env[m * hop:m * hop + n_fft] += win * win
    out = y[:, n_fft // 2:] / env[n_fft // 2:].clamp_min(1e-11)
    return out[:, :length]
