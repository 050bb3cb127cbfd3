"""Python owner of one C-ABI ``sddm_plan`` (include/sddm_b200.h): packs a reference-layout state_dict + schedule
into the library, caches per-batch workspaces (torch owns the memory) and exposes the hot-path calls on tensors."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import Config, Schedule, SCHEDULE_FIELDS, VARIANTS


def default_precision() -> int:
    v = os.environ.get("SDDM_B200_PRECISION", "bf16act").lower()
    if v in ("fp32", "0"):
        return _lib.PREC_FP32
    if v in ("bf16", "1"):
        return _lib.PREC_BF16
    if v in ("bf16act", "bf16_act", "2"):
        return _lib.PREC_BF16_ACT
    if v in ("bf16x3", "tf32", "3"):
        return _lib.PREC_BF16X3
    raise ValueError("SDDM_B200_PRECISION must be fp32, bf16, bf16act or bf16x3, got %r" % v)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Plan:
    """One plan per (weights, schedule, precision, device)."""

    def __init__(self, net_cfg: dict, weights: Dict[str, torch.Tensor], tables: Dict[str, np.ndarray], n_timestep: int,
                 precision: int, device: torch.device, pe_vector: Optional[torch.Tensor] = None):
        if device.type != "cuda":
            raise RuntimeError("sddm_b200 needs a CUDA device (no CPU fallback); got %s" % device)
        self.device = device
        self.precision = precision
        self.L = int(net_cfg["num_samples"])
        self.T = int(n_timestep)
        lib = _lib.lib()
        mults = list(net_cfg["channel_mults"])
        cfg = Config(n_timestep=n_timestep, num_samples=self.L, segment_len=net_cfg["segment_len"],
                     segment_stride=net_cfg["segment_stride"], in_channel=net_cfg["in_channel"],
                     out_channel=net_cfg["out_channel"], inner_channel=net_cfg["inner_channel"],
                     norm_groups=net_cfg["norm_groups"], n_mults=len(mults), res_blocks=net_cfg["res_blocks"],
                     precision=precision)
        if len(mults) > 8:
            raise _lib.SddmError("at most 8 channel_mults are supported")
        for i, m in enumerate(mults):
            cfg.channel_mults[i] = int(m)
        handle = C.c_void_p()
        _lib.check(lib.sddm_plan_create(C.byref(cfg), C.byref(handle)))
        self._h = handle
        try:
            with torch.cuda.device(device):
                for name, w in weights.items():
                    w = w.detach().to("cpu", torch.float32).contiguous()
                    shape = (C.c_int64 * w.dim())(*w.shape)
                    _lib.check(lib.sddm_plan_load_weight(self._h, name.encode(), C.c_void_p(w.data_ptr()), shape, w.dim()))
                if pe_vector is not None:
                    v = pe_vector.detach().to("cpu", torch.float32).contiguous()
                    shape = (C.c_int64 * 1)(v.numel())
                    _lib.check(lib.sddm_plan_load_weight(self._h, b"noise_level_mlp.0.embedding_vector",
                                                         C.c_void_p(v.data_ptr()), shape, 1))
                keep = {k: np.ascontiguousarray(tables[k], dtype=np.float32) for k in SCHEDULE_FIELDS}
                sch = Schedule(**{k: keep[k].ctypes.data_as(C.POINTER(C.c_float)) for k in SCHEDULE_FIELDS})
                _lib.check(lib.sddm_plan_set_schedule(self._h, C.byref(sch), n_timestep + 1))
                _lib.check(lib.sddm_plan_finalize(self._h))
        except Exception:
            lib.sddm_plan_destroy(self._h)
            self._h = None
            raise
        self._ws: Dict[int, torch.Tensor] = {}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().sddm_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------------------------------------
    def workspace(self, B: int) -> torch.Tensor:
        ws = self._ws.get(B)
        if ws is None:
            nbytes = int(_lib.lib().sddm_workspace_bytes(self._h, B))
            if nbytes == 0:
                raise _lib.SddmError("workspace query failed")
            self._ws = {}   # keep one workspace alive (they are large)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws[B] = ws
        return ws

    def launches_per_eps(self) -> int:
        return int(_lib.lib().sddm_plan_launches_per_eps(self._h))

    def _check_wave(self, t: torch.Tensor, name: str) -> torch.Tensor:
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor (no CPU fallback), got %s" % (name, t.device))
        if t.numel() % self.L or t.shape[-1] != self.L:
            raise ValueError("%s must be [B,1,%d], got %s" % (name, self.L, tuple(t.shape)))
        return _f32c(t)

    def eps(self, cond: torch.Tensor, x_t: torch.Tensor, noise_level: Optional[torch.Tensor] = None, t: int = 0) -> torch.Tensor:
        cond, x_t = self._check_wave(cond, "x"), self._check_wave(x_t, "y_t")
        B = cond.numel() // self.L
        nl = None
        if noise_level is not None:
            nl = _f32c(noise_level.to(self.device)).reshape(-1)
            if nl.numel() == 1 and B > 1:
                nl = nl.expand(B).contiguous()
            if nl.numel() != B:
                raise ValueError("noise level must have one entry per row")
        out = torch.empty_like(cond)
        ws = self.workspace(B)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_eps(self._h, _ptr(cond), _ptr(x_t), _ptr(nl), int(t), _ptr(out), B, _ptr(ws),
                                           ws.numel(), C.c_void_p(st)))
        return out

    def sample(self, cond: torch.Tensor, variant: str, noises: Optional[torch.Tensor] = None, seed: int = 0, row0: int = 0,
               trace: bool = False):
        """Full reverse loop.  noises: None (Philox) or [T,B,1,L]/[T,B,L].  Returns out or (out, eps_trace, x_trace)."""
        cond = self._check_wave(cond, "condition")
        B = cond.numel() // self.L
        if noises is not None:
            noises = _f32c(noises.to(self.device))
            if noises.numel() != self.T * B * self.L:
                raise ValueError("noises must hold T*B*L = %d values, got %d" % (self.T * B * self.L, noises.numel()))
        out = torch.empty_like(cond)
        eps_tr = torch.empty((self.T,) + tuple(cond.shape), device=self.device) if trace else None
        x_tr = torch.empty((self.T,) + tuple(cond.shape), device=self.device) if trace else None
        ws = self.workspace(B)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_sample(self._h, VARIANTS[variant], _ptr(cond), _ptr(noises), C.c_uint64(seed & (2 ** 64 - 1)),
                                              int(row0), _ptr(out), _ptr(eps_tr), _ptr(x_tr), B, _ptr(ws), ws.numel(),
                                              C.c_void_p(st)))
        return (out, eps_tr, x_tr) if trace else out

    def enhance_host(self, cond_host: torch.Tensor, variant: str, seed: int = 0, row0: int = 0, max_rows: int = 0,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host-buffer entry point (H2D + loop + D2H inside the library); cond_host is a CPU tensor (pin it)."""
        if cond_host.is_cuda:
            raise ValueError("enhance_host takes host tensors")
        cond_host = cond_host.detach().to(torch.float32).contiguous()
        if cond_host.shape[-1] != self.L:
            raise ValueError("condition must be [B,1,%d]" % self.L)
        B = cond_host.numel() // self.L
        if out is None:
            out = torch.empty_like(cond_host, pin_memory=cond_host.is_pinned())
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().sddm_enhance_host(self._h, VARIANTS[variant], _ptr(cond_host), _ptr(out), B,
                                                    C.c_uint64(seed & (2 ** 64 - 1)), int(row0), int(max_rows)))
        return out

    def profile(self, on: bool) -> None:
        """Enable / reset per-launch CUDA-event timing inside the library (bench.py roofline)."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().sddm_profile_enable(self._h, int(on)))

    def profile_report(self):
        """[{label, ms, launches, flops_per_row, bytes_per_row, tensor_cores}] for every op of the program."""
        lib, out = _lib.lib(), []
        with torch.cuda.device(self.device):
            for i in range(int(lib.sddm_plan_num_ops(self._h))):
                ms, n, fl, by, tc = C.c_double(), C.c_int64(), C.c_double(), C.c_double(), C.c_int()
                label = C.create_string_buffer(128)
                _lib.check(lib.sddm_profile_read(self._h, i, C.byref(ms), C.byref(n), C.byref(fl), C.byref(by), C.byref(tc),
                                                 label, 128))
                out.append(dict(label=label.value.decode(), ms=ms.value, launches=n.value, flops_per_row=fl.value,
                                bytes_per_row=by.value, tensor_cores=bool(tc.value), kernel={0: "cuda_core", 1: "conv3x3_tc_kernel", 2: "conv_row_kernel"}.get(tc.value, "?")))
        return out

    def fetch(self, node: str, B: int) -> torch.Tensor:
        """Debug: NHWC activation of a UNet node from the last eps/sample call with batch B -> [B,C,H,W]."""
        chw = (C.c_int64 * 3)()
        ws = self.workspace(B)
        lib = _lib.lib()
        _lib.check(lib.sddm_debug_fetch(self._h, node.encode(), _ptr(ws), B, None, chw, None))
        c, h, w = int(chw[0]), int(chw[1]), int(chw[2])
        buf = torch.empty((B, h, w, c), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(lib.sddm_debug_fetch(self._h, node.encode(), _ptr(ws), B, _ptr(buf), chw, C.c_void_p(st)))
        return buf.permute(0, 3, 1, 2)
