"""ctypes binding of include/sddm_b200.h.  Loading is lazy so that host-only logic (config parsing, schedule
tables, sharding) imports without the extension; every compute call goes through :func:`lib`, which fails
loudly when ``csrc/libsddm_b200.so`` has not been built (``python -m sddm_b200.build`` or
``__graft_entry__.build()``).  No fallback of any kind."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

PREC_FP32, PREC_BF16, PREC_BF16_ACT, PREC_BF16X3 = 0, 1, 2, 3
VARIANTS = {"original": 0, "condition_in": 1, "sr3": 2, "supportive": 3, "conditional": 4}

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "csrc", "libsddm_b200.so")
if os.environ.get("SDDM_B200_LIB"):   # A/B runs: load another build of the same ABI (e.g. csrc/libsddm_b200_prev.so)
    _LIB_PATH = os.path.abspath(os.environ["SDDM_B200_LIB"])
_lib: Optional[C.CDLL] = None


class SddmError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("n_timestep", C.c_int32), ("num_samples", C.c_int32), ("segment_len", C.c_int32),
                ("segment_stride", C.c_int32), ("in_channel", C.c_int32), ("out_channel", C.c_int32),
                ("inner_channel", C.c_int32), ("norm_groups", C.c_int32), ("n_mults", C.c_int32),
                ("channel_mults", C.c_int32 * 8), ("res_blocks", C.c_int32), ("precision", C.c_int32),
                ("reserved", C.c_int32 * 4)]


SCHEDULE_FIELDS = ("betas", "alphas", "sqrt_alpha_bar", "predicted_noise_coeff", "sigma", "supportive_gamma",
                   "supportive_sigma_hat", "sqrt_delta", "c_xt", "c_yt", "c_epst", "sqrt_delta_estimated")


class DwConfig(C.Structure):
    _fields_ = [("n_timestep", C.c_int32), ("freq_bins", C.c_int32), ("residual_channels", C.c_int32),
                ("residual_layers", C.c_int32), ("dilation_cycle_length", C.c_int32), ("hop_samples", C.c_int32),
                ("noise_condition", C.c_int32), ("precision", C.c_int32), ("reserved", C.c_int32 * 4)]


class WgConfig(C.Structure):
    _fields_ = [("n_timestep", C.c_int32), ("hop_samples", C.c_int32), ("noise_condition", C.c_int32), ("precision", C.c_int32),
                ("reserved", C.c_int32 * 4)]


NOISE_CONDITIONS = {"sqrt_alpha_bar": 0, "time_step": 1}


class Schedule(C.Structure):
    _fields_ = [(k, C.POINTER(C.c_float)) for k in SCHEDULE_FIELDS]


# name -> (restype, argtypes); must list every symbol declared in include/sddm_b200.h
SIGNATURES = {
    "sddm_last_error": (C.c_char_p, []),
    "sddm_version": (C.c_int, []),
    "sddm_launch_count": (C.c_uint64, []),
    "sddm_plan_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "sddm_plan_destroy": (None, [C.c_void_p]),
    "sddm_plan_load_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "sddm_plan_set_schedule": (C.c_int, [C.c_void_p, C.POINTER(Schedule), C.c_int]),
    "sddm_plan_finalize": (C.c_int, [C.c_void_p]),
    "sddm_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "sddm_eps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                           C.c_size_t, C.c_void_p]),
    "sddm_x_T": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_int,
                           C.c_void_p]),
    "sddm_p_step": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64,
                              C.c_int, C.c_int, C.c_void_p]),
    "sddm_x_T_raw": (C.c_int, [C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p,
                               C.c_int, C.c_int, C.c_void_p]),
    "sddm_p_step_raw": (C.c_int, [C.c_int, C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                  C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sddm_q_sample_raw": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "sddm_sample": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sddm_enhance_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_int]),
    "sddm_istft_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "sddm_istft": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sddm_chunk_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "sddm_regroup_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "sddm_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sddm_overlap_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sddm_stft_features": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "sddm_dw_plan_create": (C.c_int, [C.POINTER(DwConfig), C.POINTER(C.c_void_p)]),
    "sddm_dw_plan_destroy": (None, [C.c_void_p]),
    "sddm_dw_plan_load_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "sddm_dw_plan_set_schedule": (C.c_int, [C.c_void_p, C.POINTER(Schedule), C.c_int]),
    "sddm_dw_plan_finalize": (C.c_int, [C.c_void_p]),
    "sddm_dw_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "sddm_dw_condition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sddm_dw_eps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                              C.c_void_p]),
    "sddm_dw_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "sddm_dw_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "sddm_dw_profile_read": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "sddm_dw_debug_fetch": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64),
                                      C.c_void_p]),
    "sddm_wg_plan_create": (C.c_int, [C.POINTER(WgConfig), C.POINTER(C.c_void_p)]),
    "sddm_wg_plan_destroy": (None, [C.c_void_p]),
    "sddm_wg_plan_load_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "sddm_wg_plan_set_schedule": (C.c_int, [C.c_void_p, C.POINTER(Schedule), C.c_int]),
    "sddm_wg_plan_finalize": (C.c_int, [C.c_void_p]),
    "sddm_wg_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "sddm_wg_eps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                              C.c_size_t, C.c_void_p]),
    "sddm_wg_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "sddm_wg_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "sddm_wg_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "sddm_wg_debug_fetch": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64),
                                      C.c_void_p]),
    "sddm_plan_launches_per_eps": (C.c_int, [C.c_void_p]),
    "sddm_plan_num_ops": (C.c_int, [C.c_void_p]),
    "sddm_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "sddm_profile_read": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    "sddm_debug_fetch": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "sddm_var_schedule": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sddm_var_noise_level": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "sddm_var_mix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "sddm_var_posterior": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "sddm_debug_umma_probe": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "sddm_debug_tc_trace": (C.c_int, [C.c_int, C.c_void_p]),
    "sddm_debug_row_trace": (C.c_int, [C.c_int, C.c_void_p]),
    "sddm_debug_hang": (C.c_int, [C.c_int, C.c_void_p]),
    "sddm_debug_umma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
}


def library_path() -> str:
    return _LIB_PATH


def lib() -> C.CDLL:
    """The loaded extension; raises SddmError (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise SddmError(
                "sddm_b200 CUDA extension not built: %s is missing. Build it with `python -c \"import __graft_entry__ "
                "as g; g.build()\"` (needs nvcc). There is no CPU fallback." % _LIB_PATH)
        handle = C.CDLL(_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError => header / library mismatch, surfaced loudly
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().sddm_last_error()
        raise SddmError("sddm_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def launch_count() -> int:
    return int(lib().sddm_launch_count())
