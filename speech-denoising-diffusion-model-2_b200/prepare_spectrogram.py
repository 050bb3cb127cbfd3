"""STFT feature front-end — host mirror of reference prepare_spectrogram.py:13-55.

The reference builds ``TT.Spectrogram(n_fft=window_length, hop_length, window_fn=torch.hamming_window, power=1, normalized=True)``
and ``TT.MelSpectrogram(n_fft, hop_length, f_min=20, f_max=sr/2, n_mels, sample_rate, power=1, normalized=True)`` (default hann
window), applies ``clamp((log10(x) - 1 + 5) / 5, 0, 1)`` and writes ``<wav>.spec.npy`` / ``<wav>.mel.npy``.  Here the two transforms
are callables with the same constructor arguments that run ONE CUDA kernel (``sddm_stft_features``: framing, window, 1024-point FFT
with warp shuffles, magnitude, filterbank, log / clamp); the window and the HTK filterbank are built with the same torch ops
torchaudio uses, so they match bit for bit.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from glob import glob
from typing import Callable, Optional

import numpy as np
import torch

from . import _lib


def melscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk"): [n_freqs, n_mels] triangular filters."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]                                  # torchaudio _create_triangular_filterbank
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down_slopes, up_slopes))


class Spectrogram:
    """TT.Spectrogram(n_fft, hop_length, window_fn, power=1, normalized=True) on CUDA tensors [..., L] -> [..., n_fft/2+1, frames]."""

    def __init__(self, n_fft: int = 400, hop_length: Optional[int] = None, window_fn: Callable[..., torch.Tensor] = torch.hann_window,
                 power: float = 1, normalized: bool = True, log_clamp: bool = False):
        if power != 1 or not normalized:
            raise NotImplementedError("only power=1, normalized=True (the reference's settings) are implemented")
        self.n_fft = n_fft
        self.hop_length = hop_length if hop_length is not None else n_fft // 2
        self.window = window_fn(n_fft)
        self.inv_norm = float(1.0 / self.window.pow(2.0).sum().sqrt())
        self.log_clamp = log_clamp
        self._dev = {}

    def _on(self, device, name, t):
        key = (str(device), name)
        if key not in self._dev:
            self._dev[key] = t.to(device=device, dtype=torch.float32).contiguous()
        return self._dev[key]

    def _on_int(self, device, name, t):
        key = (str(device), name)
        if key not in self._dev:
            self._dev[key] = t.to(device=device, dtype=torch.int32).contiguous()
        return self._dev[key]

    def _run(self, waveform: torch.Tensor, fb: Optional[torch.Tensor], n_mels: int) -> torch.Tensor:
        if not waveform.is_cuda:
            raise RuntimeError("sddm_b200 Spectrogram needs a CUDA tensor (no CPU fallback)")
        shape = waveform.shape
        x = waveform.detach().to(torch.float32).reshape(-1, shape[-1]).contiguous()
        B, L = x.shape
        frames = 1 + L // self.hop_length
        n_out = n_mels if fb is not None else self.n_fft // 2 + 1
        out = torch.empty((B, n_out, frames), device=x.device, dtype=torch.float32)
        win = self._on(x.device, "window", self.window)
        fbd = lo = hi = None
        if fb is not None:                       # band of non-zero weights per filter: the kernel skips the zeros
            fbd = self._on(x.device, "fb", fb)
            if (str(x.device), "lo") not in self._dev:      # computed once per device: this runs on the host, ahead of every launch otherwise
                nz = fb != 0
                first = torch.where(nz.any(0), nz.float().argmax(0), torch.zeros(fb.shape[1], dtype=torch.long))
                last = torch.where(nz.any(0), fb.shape[0] - nz.flip(0).float().argmax(0), torch.zeros(fb.shape[1], dtype=torch.long))
                self._on_int(x.device, "lo", first)
                self._on_int(x.device, "hi", last)
            lo, hi = self._dev[(str(x.device), "lo")], self._dev[(str(x.device), "hi")]
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_stft_features(C.c_void_p(x.data_ptr()), B, L, self.n_fft, self.hop_length, C.c_void_p(win.data_ptr()),
                                                     C.c_float(self.inv_norm), C.c_void_p(fbd.data_ptr()) if fbd is not None else None,
                                                     C.c_void_p(lo.data_ptr()) if lo is not None else None,
                                                     C.c_void_p(hi.data_ptr()) if hi is not None else None,
                                                     n_mels, int(self.log_clamp), C.c_void_p(out.data_ptr()), C.c_void_p(st)))
        return out.reshape(tuple(shape[:-1]) + (n_out, frames))

    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        return self._run(waveform, None, 0)


class MelSpectrogram(Spectrogram):
    """TT.MelSpectrogram(sample_rate, n_fft, hop_length, f_min, f_max, n_mels, power=1, normalized=True) (hann window)."""

    def __init__(self, sample_rate: int = 16000, n_fft: int = 400, hop_length: Optional[int] = None, f_min: float = 0.0,
                 f_max: Optional[float] = None, n_mels: int = 128, power: float = 1.0, normalized: bool = True, log_clamp: bool = False):
        super().__init__(n_fft=n_fft, hop_length=hop_length, window_fn=torch.hann_window, power=power, normalized=normalized,
                         log_clamp=log_clamp)
        self.n_mels = n_mels
        self.fb = melscale_fbanks(n_fft // 2 + 1, f_min, f_max if f_max is not None else float(sample_rate // 2), n_mels, sample_rate)

    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        return self._run(waveform, self.fb, self.n_mels)


def istft(spec: torch.Tensor, n_fft: int = 1024, hop_length: int = 256, window: Optional[torch.Tensor] = None,
          length: Optional[int] = None) -> torch.Tensor:
    """Overlap-add inverse STFT on the device = torch.istft(spec, n_fft, hop_length, window=window, center=True, length=length) for a
    complex one-sided spectrogram [B, n_fft // 2 + 1, frames] (C ABI: sddm_istft).  CUDA tensors only: there is no CPU fallback."""
    import ctypes as C
    from . import _lib
    if not spec.is_cuda:
        raise RuntimeError("istft (sddm_b200) needs a CUDA tensor: there is no CPU fallback")
    if not spec.is_complex() or spec.dim() != 3 or spec.shape[1] != n_fft // 2 + 1:
        raise ValueError("spec must be complex [B, %d, frames], got %s %s" % (n_fft // 2 + 1, spec.dtype, tuple(spec.shape)))
    dev = spec.device
    B, _, frames = spec.shape
    win = (torch.hann_window(n_fft) if window is None else window).to(dev, torch.float32).contiguous()
    L = hop_length * (frames - 1) if length is None else int(length)
    ri = torch.view_as_real(spec.to(torch.complex64).contiguous()).contiguous()
    out = torch.empty((B, L), device=dev, dtype=torch.float32)
    lib = _lib.lib()
    nbytes = int(lib.sddm_istft_workspace_bytes(B, frames))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.sddm_istft(C.c_void_p(ri.data_ptr()), B, frames, n_fft, hop_length, C.c_void_p(win.data_ptr()), L,
                                  C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), nbytes, C.c_void_p(st)))
    return out


def features(audio: torch.Tensor, config) -> tuple:
    """Body of the reference's per-file loop (prepare_spectrogram.py:38-55) on a CUDA waveform [1, L] or [B, L]:
    returns (mel, spec) already log-compressed and clamped to [0, 1]."""
    window_length = config["spectrogram"]["window_length"]
    hop_samples = config["spectrogram"]["hop_samples"]
    n_mels = config["mel_spectrogram"]["n_mels"]
    sample_rate = config["sample_rate"]
    spectrogram = Spectrogram(n_fft=window_length, hop_length=hop_samples, window_fn=torch.hamming_window, power=1, normalized=True,
                              log_clamp=True)
    mel_spec = MelSpectrogram(n_fft=window_length, hop_length=hop_samples, f_min=20.0, f_max=sample_rate / 2.0, n_mels=n_mels,
                              sample_rate=sample_rate, power=1.0, normalized=True, log_clamp=True)
    return mel_spec(audio), spectrogram(audio)


def main(path: str, config, device: str = "cuda") -> int:
    """prepare_spectrogram.main: every ``*.wav`` under `path` -> ``<wav>.mel.npy`` / ``<wav>.spec.npy`` (PCM wav via scipy)."""
    from scipy.io import wavfile
    n = 0
    for filename in sorted(glob(f"{path}/**/*.wav", recursive=True)):
        sr, data = wavfile.read(filename)
        assert sr == config["sample_rate"]
        if data.dtype.kind == "i":
            data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)      # torchaudio.load(normalize=True)
        audio = torch.from_numpy(np.ascontiguousarray(data.T if data.ndim == 2 else data[None])).to(device)
        mel, spec = features(audio, config)
        np.save(f"{filename}.mel.npy", torch.squeeze(mel).cpu().numpy())
        np.save(f"{filename}.spec.npy", torch.squeeze(spec).cpu().numpy())
        n += 1
    return n
