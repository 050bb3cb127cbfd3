"""``WaveGrad`` — host mirror of reference model/wavegrad.py:140-179 (config_wavegrad.json's denoiser).

Parameter container only (same layers, names, initialisers and construction order as the reference, so ``state_dict``
and default initialisation match); ``forward`` runs the CUDA plan (``sddm_wg_*`` in include/sddm_b200.h): fp32 CUDA-core
parity path or the tcgen05 bf16 path.
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


class Conv1d(nn.Conv1d):
    """orthogonal weight, zero bias — applied by nn.Conv1d.__init__ (through the override) and once more afterwards,
    exactly as reference wavegrad.py:9-16 consumes the RNG."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.orthogonal_(self.weight)
        nn.init.zeros_(self.bias)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("sddm_b200 host modules hold parameters only; compute runs in the CUDA plan")


class PositionalEncoding(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim


class FiLM(_Holder):
    def __init__(self, input_size, output_size):
        super().__init__()
        self.encoding = PositionalEncoding(input_size)
        self.input_conv = nn.Conv1d(input_size, input_size, 3, padding=1)
        self.output_conv = nn.Conv1d(input_size, output_size * 2, 3, padding=1)
        nn.init.xavier_uniform_(self.input_conv.weight)
        nn.init.xavier_uniform_(self.output_conv.weight)
        nn.init.zeros_(self.input_conv.bias)
        nn.init.zeros_(self.output_conv.bias)


class UBlock(_Holder):
    def __init__(self, input_size, hidden_size, factor, dilation):
        super().__init__()
        assert isinstance(dilation, (list, tuple)) and len(dilation) == 4
        self.factor = factor
        self.block1 = Conv1d(input_size, hidden_size, 1)
        self.block2 = nn.ModuleList([Conv1d(input_size, hidden_size, 3, dilation=dilation[0], padding=dilation[0]),
                                     Conv1d(hidden_size, hidden_size, 3, dilation=dilation[1], padding=dilation[1])])
        self.block3 = nn.ModuleList([Conv1d(hidden_size, hidden_size, 3, dilation=dilation[2], padding=dilation[2]),
                                     Conv1d(hidden_size, hidden_size, 3, dilation=dilation[3], padding=dilation[3])])


class DBlock(_Holder):
    def __init__(self, input_size, hidden_size, factor):
        super().__init__()
        self.factor = factor
        self.residual_dense = Conv1d(input_size, hidden_size, 1)
        self.conv = nn.ModuleList([Conv1d(input_size, hidden_size, 3, dilation=1, padding=1),
                                   Conv1d(hidden_size, hidden_size, 3, dilation=2, padding=2),
                                   Conv1d(hidden_size, hidden_size, 3, dilation=4, padding=4)])


class WaveGradPlan:
    """Python owner of one C-ABI ``sddm_wg_plan``."""
    HOP, N_MELS = 300, 128

    def __init__(self, weights: Dict[str, torch.Tensor], tables: Dict[str, np.ndarray], n_timestep: int, noise_condition: str,
                 precision: int, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("sddm_b200 needs a CUDA device (no CPU fallback); got %s" % device)
        if noise_condition not in _lib.NOISE_CONDITIONS:
            raise NotImplementedError(noise_condition)
        self.device, self.T, self.precision = device, int(n_timestep), precision
        lib = _lib.lib()
        c = _lib.WgConfig(n_timestep=n_timestep, hop_samples=self.HOP, noise_condition=_lib.NOISE_CONDITIONS[noise_condition],
                          precision=precision)
        h = C.c_void_p()
        _lib.check(lib.sddm_wg_plan_create(C.byref(c), C.byref(h)))
        self._h = h
        try:
            with torch.cuda.device(device):
                for name, w in weights.items():
                    w = w.detach().to("cpu", torch.float32).contiguous()
                    shape = (C.c_int64 * w.dim())(*w.shape)
                    _lib.check(lib.sddm_wg_plan_load_weight(self._h, name.encode(), C.c_void_p(w.data_ptr()), shape, w.dim()))
                keep = {k: np.ascontiguousarray(tables[k], dtype=np.float32) for k in _lib.SCHEDULE_FIELDS}
                sch = _lib.Schedule(**{k: keep[k].ctypes.data_as(C.POINTER(C.c_float)) for k in _lib.SCHEDULE_FIELDS})
                _lib.check(lib.sddm_wg_plan_set_schedule(self._h, C.byref(sch), n_timestep + 1))
                _lib.check(lib.sddm_wg_plan_finalize(self._h))
        except Exception:
            lib.sddm_wg_plan_destroy(self._h)
            self._h = None
            raise
        self._ws, self._ws_key = None, None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().sddm_wg_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    def workspace(self, B: int, frames: int) -> torch.Tensor:
        if self._ws_key != (B, frames):
            n = int(_lib.lib().sddm_wg_workspace_bytes(self._h, B, frames))
            if n == 0:
                raise _lib.SddmError("workspace query failed")
            self._ws = None
            self._ws = torch.empty(n, dtype=torch.uint8, device=self.device)
            self._ws_key = (B, frames)
        return self._ws

    def _spec(self, spec):
        if not spec.is_cuda:
            raise RuntimeError("spectrogram must be a CUDA tensor (no CPU fallback), got %s" % spec.device)
        if spec.dim() != 3 or spec.shape[1] != self.N_MELS:
            raise ValueError("spectrogram must be [B,%d,frames], got %s" % (self.N_MELS, tuple(spec.shape)))
        return _f32c(spec)

    def eps(self, spec, audio, noise_level=None, t: int = 0):
        spec = self._spec(spec)
        B, frames = spec.shape[0], spec.shape[2]
        if not audio.is_cuda:
            raise RuntimeError("audio must be a CUDA tensor (no CPU fallback)")
        audio = _f32c(audio)
        if audio.numel() != B * self.HOP * frames:
            raise ValueError("audio must hold B x %d samples, got %s" % (self.HOP * frames, tuple(audio.shape)))
        nl = None
        if noise_level is not None:
            nl = _f32c(noise_level.to(self.device)).reshape(-1)
            if nl.numel() == 1 and B > 1:
                nl = nl.expand(B).contiguous()
            if nl.numel() != B:
                raise ValueError("noise_scale must have one entry per row")
        out = torch.empty((B, self.HOP * frames), device=self.device)
        ws = self.workspace(B, frames)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_wg_eps(self._h, _ptr(spec), _ptr(audio), _ptr(nl), int(t), _ptr(out), B, frames, _ptr(ws),
                                              ws.numel(), C.c_void_p(st)))
        return out

    def sample(self, spec, noises=None, seed: int = 0, row0: int = 0, trace: bool = False):
        spec = self._spec(spec)
        B, frames = spec.shape[0], spec.shape[2]
        Ls = self.HOP * frames
        if noises is not None:
            noises = _f32c(noises.to(self.device))
            if noises.numel() != self.T * B * Ls:
                raise ValueError("noises must hold T*B*L = %d values, got %d" % (self.T * B * Ls, noises.numel()))
        out = torch.empty((B, 1, Ls), device=self.device)
        eps_tr = torch.empty((self.T, B, 1, Ls), device=self.device) if trace else None
        ws = self.workspace(B, frames)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_wg_sample(self._h, _ptr(spec), _ptr(noises), C.c_uint64(seed & (2 ** 64 - 1)), int(row0), _ptr(out),
                                                 _ptr(eps_tr), B, frames, _ptr(ws), ws.numel(), C.c_void_p(st)))
        return (out, eps_tr) if trace else out

    def profile(self, on: bool) -> None:
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().sddm_wg_profile_enable(self._h, int(on)))

    def profile_report(self):
        """(total_ms, launches, executed_flops, algorithmic_bytes) of the tcgen05 conv launches since profile(True)."""
        ms, n, fl, by = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().sddm_wg_profile_read(self._h, C.byref(ms), C.byref(n), C.byref(fl), C.byref(by)))
        return ms.value, n.value, fl.value, by.value

    def fetch(self, what: str, B: int, frames: int) -> torch.Tensor:
        """Debug: activation 'd0'..'d4' / 'u0'..'u4' of the last eps call as [B, C, L]."""
        shape = (C.c_int64 * 2)()
        ws, lib = self.workspace(B, frames), _lib.lib()
        _lib.check(lib.sddm_wg_debug_fetch(self._h, what.encode(), _ptr(ws), B, frames, None, shape, None))
        Lx, Cx = int(shape[0]), int(shape[1])
        buf = torch.empty((B, Lx, Cx), device=self.device)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(lib.sddm_wg_debug_fetch(self._h, what.encode(), _ptr(ws), B, frames, _ptr(buf), shape, C.c_void_p(st)))
        return buf.permute(0, 2, 1)


class WaveGrad(nn.Module):
    def __init__(self, **unused):
        super().__init__()
        self.downsample = nn.ModuleList([Conv1d(1, 32, 5, padding=2), DBlock(32, 128, 2), DBlock(128, 128, 2), DBlock(128, 256, 3),
                                         DBlock(256, 512, 5)])
        self.film = nn.ModuleList([FiLM(32, 128), FiLM(128, 128), FiLM(128, 256), FiLM(256, 512), FiLM(512, 512)])
        self.upsample = nn.ModuleList([UBlock(768, 512, 5, [1, 2, 1, 2]), UBlock(512, 512, 5, [1, 2, 1, 2]), UBlock(512, 256, 3, [1, 2, 4, 8]),
                                       UBlock(256, 128, 2, [1, 2, 4, 8]), UBlock(128, 128, 2, [1, 2, 4, 8])])
        self.first_conv = Conv1d(128, 768, 3, padding=1)
        self.last_conv = Conv1d(128, 1, 3, padding=1)
        self.precision: Optional[int] = None     # None -> SDDM_B200_PRECISION env / default
        self._plans: Dict[tuple, WaveGradPlan] = {}

    def _param_version(self):
        return tuple(int(p._version) for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())

    def _apply(self, fn, *a, **k):
        self._plans = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._plans = {}
        return super().load_state_dict(*a, **k)

    def get_plan(self, diffusion=None, noise_condition: str = "sqrt_alpha_bar", precision: Optional[int] = None) -> WaveGradPlan:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("WaveGrad (sddm_b200) must live on a CUDA device: call .to('cuda') first; there is no CPU fallback")
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
            for i, f in enumerate(self.film):            # PositionalEncoding frequencies, computed as the reference does (:45-46)
                count = f.encoding.dim // 2
                step = torch.arange(count, dtype=torch.float32) / count
                weights["film.%d.encoding.frequencies" % i] = torch.exp(-np.log(1e4) * step)
            plan = WaveGradPlan(weights, tables, T, noise_condition, prec, dev)
            self._plans[key] = plan
        return plan

    @torch.no_grad()
    def forward(self, spectrogram, audio, noise_scale):
        """eps_hat from spectrogram [B,128,F], audio [B,300 F], noise_scale [B] or [B,1,1] (reference :167-179, including the
        final torch.squeeze)."""
        return torch.squeeze(self.get_plan().eps(spectrogram, audio, noise_level=noise_scale))
