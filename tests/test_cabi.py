"""The C-ABI library loads, exports every symbol include/sddm_b200.h declares, and validates its arguments.
No compute call is made here (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT, UNET_CFG


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sddm_b200.h")).read()
    return sorted(set(re.findall(r"SDDM_API\s+[\w\s\*]+?\b(sddm_\w+)\s*\(", text)))


def test_header_symbols_exported_and_bound(built_lib):
    from sddm_b200 import _lib
    names = header_symbols()
    assert len(names) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.library_path()], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (sddm_\w+)", out))
    assert set(names) <= exported, set(names) - exported
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    # nothing but the C ABI leaks out of the library
    assert all(s.startswith("sddm_") for s in re.findall(r" T (\w+)", out))
    assert built_lib.sddm_version() >= 100


def _cfg(**over):
    from sddm_b200._lib import Config
    d = dict(n_timestep=100, num_samples=16448, segment_len=128, segment_stride=64, in_channel=2, out_channel=1,
             inner_channel=32, norm_groups=32, n_mults=5, res_blocks=1, precision=0)
    d.update(over)
    mults = d.pop("channel_mults", (1, 2, 3, 4, 5))
    c = Config(**d)
    for i, m in enumerate(mults):
        c.channel_mults[i] = m
    return c


def test_plan_create_validation(built_lib):
    lib = built_lib
    h = C.c_void_p()
    assert lib.sddm_plan_create(C.byref(_cfg()), C.byref(h)) == 0 and h.value
    # weights: unknown name, wrong shape, good one
    w = torch.zeros(32, 2, 3, 3)
    shape = (C.c_int64 * 4)(32, 2, 3, 3)
    assert lib.sddm_plan_load_weight(h, b"downs.0.weight", C.c_void_p(w.data_ptr()), shape, 4) == 0
    assert lib.sddm_plan_load_weight(h, b"nope.weight", C.c_void_p(w.data_ptr()), shape, 4) == -1
    assert b"unexpected weight" in lib.sddm_last_error()
    bad = (C.c_int64 * 4)(32, 3, 3, 3)
    assert lib.sddm_plan_load_weight(h, b"downs.0.weight", C.c_void_p(w.data_ptr()), bad, 4) == -1
    assert b"shape mismatch" in lib.sddm_last_error()
    # call-order errors
    assert lib.sddm_plan_finalize(h) == -2 and b"schedule" in lib.sddm_last_error()
    assert lib.sddm_workspace_bytes(h, 4) == 0
    assert lib.sddm_eps(h, None, None, None, 1, None, 1, None, 0, None) == -2
    lib.sddm_plan_destroy(h)
    # config errors mirror the reference's own checks
    for over, msg in ((dict(num_samples=16449), b"segment_stride"),            # UNetModified2.py:13 assert
                      (dict(in_channel=3), b"in_channel"),
                      (dict(precision=7), b"precision"),
                      (dict(n_mults=0), b"n_mults"),
                      (dict(num_samples=128 + 64 * 9), b"tile")):               # 10 frames: does not tile through 5 levels
        h2 = C.c_void_p()
        assert lib.sddm_plan_create(C.byref(_cfg(**over)), C.byref(h2)) == -1, over
        assert msg in lib.sddm_last_error(), (over, lib.sddm_last_error())


def test_framing_argument_checks(built_lib):
    lib = built_lib
    x = torch.zeros(16)
    p = C.c_void_p(x.data_ptr())
    assert lib.sddm_frames(p, p, 1, 11, 4, 2, None) == -1     # (n - F) % stride != 0, UNetModified2.py:13
    assert lib.sddm_overlap_add(p, p, 0, 10, 4, 2, None) == -1
    k8 = (C.c_float * 8)()
    assert lib.sddm_p_step_raw(9, k8, p, p, None, None, 0, 0, 1, 4, 1, 16, None) == -1   # unknown variant
    assert lib.sddm_p_step_raw(0, k8, p, p, None, None, 0, 0, 5, 4, 1, 16, None) == -1   # t out of range
    assert lib.sddm_p_step_raw(3, k8, p, p, None, None, 0, 0, 1, 4, 1, 16, None) == -1   # supportive needs condition


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    from sddm_b200 import SddmError
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM
    from sddm_b200.model.network import UNetModified2
    from sddm_b200.plan import Plan
    net = UNetModified2(**UNET_CFG)
    d = GaussianDiffusion("linear", 100, 1e-6, 1e-3, device="cpu")
    m = SDDM(d, net, p_transition="condition_in")
    x = torch.zeros(1, 1, 16448)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.infer(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x, x, torch.ones(1, 1, 1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.p_transition(x, 3, x)
    with pytest.raises((SddmError, RuntimeError)):
        Plan(net.cfg, dict(net.state_dict()), d.host_tables(), 100, 0, torch.device("cuda"))


def test_diffwave_and_wavegrad_plan_validation(built_lib):
    """sddm_dw_* / sddm_wg_* argument and call-order checks (no device work)."""
    from sddm_b200._lib import DwConfig, WgConfig
    lib = built_lib
    good = dict(n_timestep=200, freq_bins=513, residual_channels=64, residual_layers=30, dilation_cycle_length=10, hop_samples=256,
                noise_condition=1, precision=0)
    h = C.c_void_p()
    assert lib.sddm_dw_plan_create(C.byref(DwConfig(**good)), C.byref(h)) == 0 and h.value
    w = torch.zeros(128, 64, 3)
    assert lib.sddm_dw_plan_load_weight(h, b"residual_layers.29.dilated_conv.weight", C.c_void_p(w.data_ptr()), (C.c_int64 * 3)(128, 64, 3), 3) == 0
    assert lib.sddm_dw_plan_load_weight(h, b"residual_layers.30.dilated_conv.weight", C.c_void_p(w.data_ptr()), (C.c_int64 * 3)(128, 64, 3), 3) == -1
    assert lib.sddm_dw_plan_load_weight(h, b"residual_layers.0.dilated_conv.weight", C.c_void_p(w.data_ptr()), (C.c_int64 * 3)(128, 64, 5), 3) == -1
    assert lib.sddm_dw_plan_finalize(h) == -2 and b"schedule" in lib.sddm_last_error()
    assert lib.sddm_dw_eps(h, None, None, 1, None, 1, 4, None, 0, None) == -2
    lib.sddm_dw_plan_destroy(h)
    for over, msg in ((dict(residual_channels=128), b"residual_channels"), (dict(hop_samples=300), b"hop_samples"),
                      (dict(precision=2), b"precision"), (dict(noise_condition=5), b"noise_condition"), (dict(residual_layers=0), b"residual_layers")):
        h2 = C.c_void_p()
        assert lib.sddm_dw_plan_create(C.byref(DwConfig(**{**good, **over})), C.byref(h2)) == -1, over
        assert msg in lib.sddm_last_error(), (over, lib.sddm_last_error())
    wg = dict(n_timestep=1000, hop_samples=300, noise_condition=0, precision=0)
    h = C.c_void_p()
    assert lib.sddm_wg_plan_create(C.byref(WgConfig(**wg)), C.byref(h)) == 0 and h.value
    w = torch.zeros(512, 768, 3)
    assert lib.sddm_wg_plan_load_weight(h, b"upsample.0.block2.0.weight", C.c_void_p(w.data_ptr()), (C.c_int64 * 3)(512, 768, 3), 3) == 0
    assert lib.sddm_wg_plan_load_weight(h, b"upsample.0.block2.0.weight", C.c_void_p(w.data_ptr()), (C.c_int64 * 3)(512, 512, 3), 3) == -1
    assert lib.sddm_wg_plan_load_weight(h, b"upsample.5.block1.weight", C.c_void_p(w.data_ptr()), (C.c_int64 * 3)(512, 768, 3), 3) == -1
    assert lib.sddm_wg_plan_finalize(h) == -2
    assert lib.sddm_wg_workspace_bytes(h, 2, 4) > 0                    # sizing needs no device
    assert lib.sddm_wg_eps(h, None, None, None, 1, None, 1, 4, None, 0, None) == -2
    lib.sddm_wg_plan_destroy(h)
    for over, msg in ((dict(hop_samples=256), b"hop_samples"), (dict(precision=2), b"precision"), (dict(n_timestep=0), b"n_timestep")):
        h2 = C.c_void_p()
        assert lib.sddm_wg_plan_create(C.byref(WgConfig(**{**wg, **over})), C.byref(h2)) == -1, over
        assert msg in lib.sddm_last_error(), (over, lib.sddm_last_error())
