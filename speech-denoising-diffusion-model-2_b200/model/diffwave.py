"""``DiffWave`` — host mirror of reference model/diffwave.py:111-155 (config_diffwave.json's denoiser).

Like ``UNetModified2`` here, the module is a *parameter container*: it creates the same torch layers, in the same
order, under the same attribute names and with the same initialisers as the reference, so ``state_dict()`` keys /
shapes match (reference checkpoints load unchanged) and default initialisation under a given ``torch.manual_seed``
reproduces the reference's weights.  No torch compute: ``forward`` goes to the CUDA plan (``sddm_dw_*`` in
include/sddm_b200.h).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch
from torch import nn

from .. import _lib
from ..plan import _f32c, _ptr, default_precision


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("sddm_b200 host modules hold parameters only; compute runs in the CUDA plan")


def Conv1d(*args, **kwargs):
    """nn.Conv1d re-initialised with kaiming_normal_ (reference diffwave.py:11-14: same RNG consumption)."""
    layer = nn.Conv1d(*args, **kwargs)
    nn.init.kaiming_normal_(layer.weight)
    return layer


class DiffusionEmbedding(_Holder):
    def __init__(self, dim=128):
        super().__init__()
        self.dim = dim
        step = torch.arange(self.dim // 2) / (self.dim // 2)
        self.embedding_vector = 10.0 ** (step * 4.0 / 63)          # plain attribute, as in the reference (:28)
        self.projection1 = nn.Linear(128, 512)
        self.projection2 = nn.Linear(512, 512)


class SpectrogramUpsampler(_Holder):
    def __init__(self, freq_bins):
        super().__init__()
        self.conv1 = nn.ConvTranspose2d(1, 1, [3, 32], stride=[1, 16], padding=[1, 8])
        self.conv2 = nn.ConvTranspose2d(1, 1, [3, 32], stride=[1, 16], padding=[1, 8])


class ResidualBlock(_Holder):
    def __init__(self, freq_bins, residual_channels, dilation):
        super().__init__()
        self.dilated_conv = Conv1d(residual_channels, 2 * residual_channels, 3, padding=dilation, dilation=dilation)
        self.diffusion_projection = nn.Linear(512, residual_channels)
        self.conditioner_projection = Conv1d(freq_bins, 2 * residual_channels, 1)
        self.output_projection = Conv1d(residual_channels, residual_channels, 1)      # split=True (the reference default)
        self.output_residual = Conv1d(residual_channels, residual_channels, 1)


class DiffWavePlan:
    """Python owner of one C-ABI ``sddm_dw_plan``."""

    def __init__(self, cfg: dict, weights: Dict[str, torch.Tensor], tables: Dict[str, np.ndarray], n_timestep: int,
                 noise_condition: str, precision: int, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("sddm_b200 needs a CUDA device (no CPU fallback); got %s" % device)
        if noise_condition not in _lib.NOISE_CONDITIONS:
            raise NotImplementedError(noise_condition)
        self.device, self.precision, self.T = device, precision, int(n_timestep)
        self.freq_bins, self.hop = int(cfg["freq_bins"]), 256
        lib = _lib.lib()
        c = _lib.DwConfig(n_timestep=n_timestep, freq_bins=cfg["freq_bins"], residual_channels=cfg["residual_channels"],
                          residual_layers=cfg["residual_layers"], dilation_cycle_length=cfg["dilation_cycle_length"],
                          hop_samples=self.hop, noise_condition=_lib.NOISE_CONDITIONS[noise_condition], precision=precision)
        h = C.c_void_p()
        _lib.check(lib.sddm_dw_plan_create(C.byref(c), C.byref(h)))
        self._h = h
        try:
            with torch.cuda.device(device):
                for name, w in weights.items():
                    w = w.detach().to("cpu", torch.float32).contiguous()
                    shape = (C.c_int64 * w.dim())(*w.shape)
                    _lib.check(lib.sddm_dw_plan_load_weight(self._h, name.encode(), C.c_void_p(w.data_ptr()), shape, w.dim()))
                keep = {k: np.ascontiguousarray(tables[k], dtype=np.float32) for k in _lib.SCHEDULE_FIELDS}
                sch = _lib.Schedule(**{k: keep[k].ctypes.data_as(C.POINTER(C.c_float)) for k in _lib.SCHEDULE_FIELDS})
                _lib.check(lib.sddm_dw_plan_set_schedule(self._h, C.byref(sch), n_timestep + 1))
                _lib.check(lib.sddm_dw_plan_finalize(self._h))
        except Exception:
            lib.sddm_dw_plan_destroy(self._h)
            self._h = None
            raise
        self._ws = None
        self._ws_key = None
        self._cond_key = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().sddm_dw_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    def workspace(self, B: int, frames: int) -> torch.Tensor:
        if self._ws_key != (B, frames):
            n = int(_lib.lib().sddm_dw_workspace_bytes(self._h, B, frames))
            if n == 0:
                raise _lib.SddmError("workspace query failed")
            self._ws = None
            self._ws = torch.empty(n, dtype=torch.uint8, device=self.device)
            self._ws_key, self._cond_key = (B, frames), None
        return self._ws

    def _spec(self, spec: torch.Tensor) -> torch.Tensor:
        if not spec.is_cuda:
            raise RuntimeError("spectrogram must be a CUDA tensor (no CPU fallback), got %s" % spec.device)
        if spec.dim() != 3 or spec.shape[1] != self.freq_bins:
            raise ValueError("spectrogram must be [B,%d,frames], got %s" % (self.freq_bins, tuple(spec.shape)))
        return _f32c(spec)

    def condition(self, spec: torch.Tensor) -> None:
        """Upsampler + conditioner projections of every layer, cached in the workspace (step independent)."""
        spec = self._spec(spec)
        B, frames = spec.shape[0], spec.shape[2]
        ws = self.workspace(B, frames)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_dw_condition(self._h, _ptr(spec), B, frames, _ptr(ws), ws.numel(), C.c_void_p(st)))
        self._cond_key = (B, frames)

    def eps(self, spec: torch.Tensor, audio: torch.Tensor, diffusion_step: Optional[torch.Tensor] = None, t: int = 0,
            reuse_condition: bool = False) -> torch.Tensor:
        """eps_hat of one step.  The conditioner (upsampler + 30 projections) is recomputed from `spec` on every call unless the
        caller opts in with reuse_condition=True after a `condition(spec)` / `eps(spec, ...)` call on the SAME spectrogram batch:
        tensor identity (address, version counter) cannot tell two batches apart - the caching allocator hands a fresh batch the
        address of the one it just freed."""
        spec = self._spec(spec)
        B, frames = spec.shape[0], spec.shape[2]
        if not audio.is_cuda:
            raise RuntimeError("audio must be a CUDA tensor (no CPU fallback)")
        audio = _f32c(audio)
        if audio.numel() != B * self.hop * frames:
            raise ValueError("audio must be [B,1,%d], got %s" % (self.hop * frames, tuple(audio.shape)))
        if not (reuse_condition and self._cond_key == (B, frames)):
            self.condition(spec)
        step = None
        if diffusion_step is not None:
            step = _f32c(diffusion_step.to(self.device)).reshape(-1)
            if step.numel() == 1 and B > 1:
                step = step.expand(B).contiguous()
            if step.numel() != B:
                raise ValueError("diffusion_step must have one entry per row")
        out = torch.empty_like(audio)
        ws = self.workspace(B, frames)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_dw_eps(self._h, _ptr(audio), _ptr(step), int(t), _ptr(out), B, frames, _ptr(ws), ws.numel(),
                                              C.c_void_p(st)))
        return out

    def sample(self, spec: torch.Tensor, noises: Optional[torch.Tensor] = None, seed: int = 0, row0: int = 0, trace: bool = False):
        spec = self._spec(spec)
        B, frames = spec.shape[0], spec.shape[2]
        Ls = self.hop * frames
        if noises is not None:
            noises = _f32c(noises.to(self.device))
            if noises.numel() != self.T * B * Ls:
                raise ValueError("noises must hold T*B*L = %d values, got %d" % (self.T * B * Ls, noises.numel()))
        out = torch.empty((B, 1, Ls), device=self.device)
        eps_tr = torch.empty((self.T, B, 1, Ls), device=self.device) if trace else None
        ws = self.workspace(B, frames)
        self._cond_key = None
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_dw_sample(self._h, _ptr(spec), _ptr(noises), C.c_uint64(seed & (2 ** 64 - 1)), int(row0), _ptr(out),
                                                 _ptr(eps_tr), B, frames, _ptr(ws), ws.numel(), C.c_void_p(st)))
        self._cond_key = (B, frames)
        return (out, eps_tr) if trace else out

    def profile(self, on: bool) -> None:
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().sddm_dw_profile_enable(self._h, int(on)))

    def profile_report(self):
        """{'layer': (total_ms, launches), 'head': (...)} of the tcgen05 kernels since profile(True)."""
        out = {}
        with torch.cuda.device(self.device):
            for kind, name in ((0, "layer"), (1, "head")):
                ms, n = C.c_double(), C.c_int64()
                _lib.check(_lib.lib().sddm_dw_profile_read(self._h, kind, C.byref(ms), C.byref(n)))
                out[name] = (ms.value, n.value)
        return out

    def fetch(self, what: str, B: int, frames: int) -> torch.Tensor:
        """Debug: 'upsampled' [T,F] (last utterance), 'x' / 'skip' [B,T,64], 'cond<i>' [B,T,128] as fp32."""
        n = C.c_int64()
        ws, lib = self.workspace(B, frames), _lib.lib()
        _lib.check(lib.sddm_dw_debug_fetch(self._h, what.encode(), _ptr(ws), B, frames, None, C.byref(n), None))
        buf = torch.empty(int(n.value), device=self.device)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(lib.sddm_dw_debug_fetch(self._h, what.encode(), _ptr(ws), B, frames, _ptr(buf), C.byref(n), C.c_void_p(st)))
        return buf


class DiffWave(nn.Module):
    def __init__(self, num_samples=-1, num_timesteps=200, freq_bins=513, residual_channels=64, residual_layers=30,
                 dilation_cycle_length=10):
        super().__init__()
        self.cfg = dict(freq_bins=freq_bins, residual_channels=residual_channels, residual_layers=residual_layers,
                        dilation_cycle_length=dilation_cycle_length)
        self.input_projection = Conv1d(1, residual_channels, 1)
        self.diffusion_embedding = DiffusionEmbedding()
        self.spectrogram_upsampler = SpectrogramUpsampler(freq_bins)
        self.residual_layers = nn.ModuleList([
            ResidualBlock(freq_bins, residual_channels, 2 ** (i % dilation_cycle_length)) for i in range(residual_layers)])
        self.skip_projection = Conv1d(residual_channels, residual_channels, 1)
        self.output_projection = Conv1d(residual_channels, 1, 1)
        nn.init.zeros_(self.output_projection.weight)
        self.precision: Optional[int] = None     # None -> SDDM_B200_PRECISION env / default
        self._plans: Dict[tuple, DiffWavePlan] = {}

    def _param_version(self):
        return tuple(int(p._version) for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())

    def _apply(self, fn, *a, **k):
        self._plans = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._plans = {}
        return super().load_state_dict(*a, **k)

    def get_plan(self, diffusion=None, noise_condition: str = "time_step", precision: Optional[int] = None) -> DiffWavePlan:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("DiffWave (sddm_b200) must live on a CUDA device: call .to('cuda') first; there is no CPU fallback")
        prec = precision if precision is not None else (self.precision if self.precision is not None else default_precision())
        prec = _lib.PREC_FP32 if prec in (_lib.PREC_FP32, _lib.PREC_BF16X3) else _lib.PREC_BF16      # bf16act == bf16, bf16x3 -> fp32 for this denoiser
        tables = diffusion.host_tables() if diffusion is not None else None
        key = (diffusion.tables_key() if diffusion is not None else None, noise_condition, prec, str(dev), self._param_version())
        plan = self._plans.get(key)
        if plan is None:
            self._plans = {k: v for k, v in self._plans.items() if k[4] == key[4]}
            if diffusion is not None:
                T = diffusion.num_timesteps
            else:
                T, tables = 1, {k: np.ones(2, dtype=np.float32) for k in _lib.SCHEDULE_FIELDS}
            weights = dict(self.state_dict())
            weights["diffusion_embedding.embedding_vector"] = self.diffusion_embedding.embedding_vector
            plan = DiffWavePlan(self.cfg, weights, tables, T, noise_condition, prec, dev)
            self._plans[key] = plan
        return plan

    @torch.no_grad()
    def forward(self, spectrogram, audio, diffusion_step):
        """eps_hat [B,1,T] from spectrogram [B,F,frames], audio [B,1,T], diffusion_step [B,1,1] (reference :133-155)."""
        return self.get_plan().eps(spectrogram, audio, diffusion_step=diffusion_step).reshape(audio.shape)
