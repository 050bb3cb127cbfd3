"""Parity of the CUDA hot path (through the C ABI) against the CPU oracle and the reference-generated goldens.

Bars (BASELINE.json north_star): per-step eps_hat error max|d|/max|ref| <= 1e-3 in the high-precision mode (our fp32
CUDA-core mode, which replaces 'TF32 mode') and <= 2e-2 in bf16 (tcgen05) mode; final waveform SI-SNR >= 40 dB.
A report of every measured error is appended to gpurun_out/parity_report.txt.
"""
import os
import sys
import time

import pytest
import torch

from conftest import CFG2_GOLDEN_ROWS, L, ROOT, UNET_CFG, cfg1_condition, cfg2_inputs, rel_err, seed0_state_dict

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sddm_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

EPS_BAR = {"fp32": 1e-3, "bf16x3": 1e-3, "bf16": 2e-2, "bf16act": 2e-2}   # bf16x3 = the tensor-core high-precision ("TF32-class") mode
SNR_BAR = 40.0


def report(line):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.txt"), "a") as f:
        f.write(line + "\n")
    print(line)


@pytest.fixture(scope="module")
def dev(built_lib):
    return torch.device("cuda:0")


def prec_id(name):
    from sddm_b200 import PREC_BF16, PREC_BF16_ACT, PREC_BF16X3, PREC_FP32
    return {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16act": PREC_BF16_ACT, "bf16x3": PREC_BF16X3}[name]


def make_model(dev, T=100, start=1e-6, end=1e-3, variant="condition_in", sd=None):
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM
    from sddm_b200.model.network import UNetModified2
    if sd is None:
        sd, net = seed0_state_dict()
    else:
        net = UNetModified2(**UNET_CFG)
        net.load_state_dict({k[len("noise_estimate_model."):]: v for k, v in sd.items()})
    diff = GaussianDiffusion("linear", T, start, end, device=dev)
    return SDDM(diff, net, p_transition=variant).to(dev).eval(), sd


# ----------------------------------------------------------------------------------------------------
# elementwise kernels: bit-exact
# ----------------------------------------------------------------------------------------------------
def test_posterior_and_x_T_bit_exact(dev, golden):
    from sddm_b200.model.diffusion import GaussianDiffusion
    g = {k: v.to(dev) for k, v in golden("steps.npz").items()}
    d = GaussianDiffusion("linear", 100, 1e-6, 1e-3, device=dev)
    for t in (1, 2, 50, 100):
        outs = {"original": d.p_transition(g["x"], t, g["eps"], noise=g["z"]),
                "sr3": d.p_transition_sr3(g["x"], t, g["eps"], noise=g["z"]),
                "supportive": d.p_transition_supportive(g["x"], t, g["eps"], g["cond"], noise=g["z"]),
                "conditional": d.p_transition_conditional(g["x"], t, g["eps"], g["cond"], noise=g["z"])}
        for variant, out in outs.items():
            assert torch.equal(out, g[f"{variant}.t{t}"]), (variant, t, float((out - g[f"{variant}.t{t}"]).abs().max()))
    assert torch.equal(d.get_x_T(g["cond"], noise=g["z"]), g["get_x_T"])
    assert torch.equal(d.get_x_T_conditional(g["cond"], noise=g["z"]), g["get_x_T_conditional"])
    report("posterior/x_T kernels: bit-exact vs reference for 4 variants x t in {1,2,50,100}")


def test_frames_and_overlap_add(dev, meta, built_lib):
    import ctypes as C
    lib = built_lib
    toy = meta["framing_toy"]
    sig = torch.tensor(toy["signal"], device=dev).reshape(1, 1, 10)
    fr = torch.empty(1, 1, 4, 4, device=dev)
    assert lib.sddm_frames(C.c_void_p(sig.data_ptr()), C.c_void_p(fr.data_ptr()), 1, 10, 4, 2, None) == 0
    back = torch.empty(1, 1, 10, device=dev)
    assert lib.sddm_overlap_add(C.c_void_p(fr.data_ptr()), C.c_void_p(back.data_ptr()), 1, 10, 4, 2, None) == 0
    torch.cuda.synchronize()
    assert fr[0, 0].tolist() == toy["frames"] and back.flatten().tolist() == toy["overlap_add"]
    x = torch.randn(3, 1, L, generator=torch.Generator().manual_seed(2))
    fr = torch.empty(3, 1, 256, 128, device=dev)
    xd = x.to(dev)
    assert lib.sddm_frames(C.c_void_p(xd.data_ptr()), C.c_void_p(fr.data_ptr()), 3, L, 128, 64, None) == 0
    back = torch.empty(3, 1, L, device=dev)
    assert lib.sddm_overlap_add(C.c_void_p(fr.data_ptr()), C.c_void_p(back.data_ptr()), 3, L, 128, 64, None) == 0
    torch.cuda.synchronize()
    ref_fr = O.signal_to_frames(x, 128, 64)
    assert torch.equal(fr.cpu(), ref_fr) and torch.equal(back.cpu(), O.overlap_add(ref_fr, L, 64))


def test_philox_noise(dev):
    """Pure-noise x_T (variant 'original'): N(0,1) moments, determinism in the seed, independence of rows/seeds/batching."""
    from sddm_b200.model.diffusion import GaussianDiffusion
    d = GaussianDiffusion("linear", 10, device=dev)
    like = torch.zeros(8, 1, L, device=dev)
    a = d._x_T("original", like, None, 123)
    b = d._x_T("original", like, None, 123)
    c = d._x_T("original", like, None, 124)
    assert torch.equal(a, b) and not torch.equal(a, c)
    v = a.double().flatten()
    assert abs(float(v.mean())) < 0.01 and abs(float(v.var()) - 1.0) < 0.01
    assert abs(float((v ** 4).mean()) - 3.0) < 0.1 and abs(float((v ** 3).mean())) < 0.05
    assert float(v.abs().max()) < 7.0 and torch.isfinite(v).all()
    rows = a.reshape(8, -1).double()
    corr = torch.corrcoef(rows)
    assert float((corr - torch.eye(8, device=dev)).abs().max()) < 0.05


# ----------------------------------------------------------------------------------------------------
# tcgen05 descriptor self-test: one 128 x N x K bf16 MMA tile through smem descriptors / TMEM, vs a host fp64 reference
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_umma_probe(dev, built_lib, variant):
    """variant 0: canonical K-major no-swizzle operand; 1: the stride-1 halo-window geometry of conv_tc.cu (SBO 160 B,
    LBO 181*16 B, start address offset by the centre tap); 2: the stride-2 (parity-split) geometry."""
    import ctypes as C
    from sddm_b200 import _lib
    for N, K in ((32, 16), (96, 64), (160, 32), (256, 64)):
        err = C.c_float(-1.0)
        _lib.check(built_lib.sddm_debug_umma_probe(variant, N, K, C.byref(err)))
        report(f"umma probe variant={variant} N={N} K={K}: max abs err {err.value:.3e}")
        assert 0.0 <= err.value < 1e-3, (variant, N, K, err.value)


# ----------------------------------------------------------------------------------------------------
# the denoiser, node by node
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["fp32", "bf16x3", "bf16", "bf16act"])
def test_unet_nodes_vs_oracle(dev, prec):
    """Random weights with non-trivial GroupNorm affine terms; every UNet node compared with the oracle."""
    cfg = dict(UNET_CFG)
    sd = O.random_state_dict(cfg, seed=3)
    model, _ = make_model(dev, sd=sd)
    net = model.noise_estimate_model
    net.precision = prec_id(prec)
    g = torch.Generator().manual_seed(21)
    x = (0.1 * torch.randn(2, 1, L, generator=g)).clamp(-1, 1)
    y = (0.5 * torch.randn(2, 1, L, generator=g)).clamp(-1, 1)
    nl = torch.tensor([0.97, 0.9995]).reshape(2, 1, 1)
    taps = {}
    ref = O.unet_forward(sd, cfg, x, y, nl, taps=taps)
    out = net(x.to(dev), y.to(dev), nl.to(dev)).cpu()
    plan = net.get_plan()
    worst = 0.0
    for name, t in taps.items():
        got = plan.fetch("frames" if name == "final_conv" else name, 2).cpu()
        e = rel_err(got, t)
        worst = max(worst, e)
        report(f"node[{prec}] {name:12s} shape={tuple(t.shape)} rel_err={e:.3e}")
    e = rel_err(out, ref)
    report(f"node[{prec}] eps_hat rel_err={e:.3e} (bar {EPS_BAR[prec]:.0e}), worst node {worst:.3e}")
    assert e <= EPS_BAR[prec]
    assert worst <= (1e-3 if prec in ("fp32", "bf16x3") else 3e-2)


@pytest.mark.parametrize("prec", ["fp32", "bf16x3", "bf16", "bf16act"])
@pytest.mark.parametrize("shape", ["two_levels_two_resblocks", "three_levels_wide", "batch_of_one", "odd_batch"])
def test_other_unet_configs_vs_oracle(dev, prec, shape):
    """Shapes other than config_unet.json (different depth, res_blocks, channel widths, frame grid, batch sizes 1 / 3 / 5):
    the op program, tile masks, ring plans and weight residency are all derived from the config."""
    from sddm_b200.model.network import UNetModified2
    if shape == "two_levels_two_resblocks":
        cfg = dict(num_samples=64 + 31 * 32, in_channel=2, out_channel=1, inner_channel=32, norm_groups=32, channel_mults=(1, 2),
                   res_blocks=2, dropout=0, segment_len=64, segment_stride=32)
        B = 3
    elif shape == "three_levels_wide":
        cfg = dict(num_samples=128 + 63 * 64, in_channel=2, out_channel=1, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4),
                   res_blocks=1, dropout=0, segment_len=128, segment_stride=64)
        B = 5
    elif shape == "odd_batch":   # the 64-wide level tiles two samples per MMA: the last pair is half empty
        cfg = dict(UNET_CFG)
        B = 5
    else:
        cfg = dict(UNET_CFG)
        B = 1
    sd = O.random_state_dict(cfg, seed=11)
    net = UNetModified2(**cfg)
    net.load_state_dict({k[len("noise_estimate_model."):]: v for k, v in sd.items()})
    net = net.to(dev).eval()
    net.precision = prec_id(prec)
    Lc = cfg["num_samples"]
    g = torch.Generator().manual_seed(5)
    x = (0.1 * torch.randn(B, 1, Lc, generator=g)).clamp(-1, 1)
    y = (0.4 * torch.randn(B, 1, Lc, generator=g)).clamp(-1, 1)
    nl = torch.linspace(0.95, 0.9999, B).reshape(B, 1, 1)
    ref = O.unet_forward(sd, cfg, x, y, nl)
    out = net(x.to(dev), y.to(dev), nl.to(dev)).cpu()
    e = rel_err(out, ref)
    report(f"config[{shape}][{prec}] B={B}: eps_hat rel_err={e:.3e} (bar {EPS_BAR[prec]:.0e})")
    assert e <= EPS_BAR[prec]


@pytest.mark.parametrize("prec", ["fp32", "bf16x3", "bf16", "bf16act"])
def test_unet_eps_vs_reference_golden(dev, golden, meta, prec):
    model, _ = make_model(dev)
    net = model.noise_estimate_model
    net.precision = prec_id(prec)
    g = torch.Generator().manual_seed(meta["eps_inputs"]["seed"])
    x = (0.1 * torch.randn(2, 1, L, generator=g)).clamp(-1, 1)
    y = (0.3 * torch.randn(2, 1, L, generator=g)).clamp(-1, 1)
    nl = torch.tensor(meta["eps_inputs"]["noise_level"]).reshape(2, 1, 1)      # per-row noise levels
    out = net(x.to(dev), y.to(dev), nl.to(dev)).cpu()
    e = rel_err(out, golden("unet_eps.npz")["eps"])
    report(f"eps_hat[{prec}] vs reference golden: rel_err={e:.3e} (bar {EPS_BAR[prec]:.0e})")
    assert e <= EPS_BAR[prec]
    # table path (row t of the precomputed embedding table) == explicit noise level sqrt_alpha_bar[t]
    plan = net.get_plan(model.diffusion)
    lvl = model.diffusion.sqrt_alpha_bar[37] * torch.ones(2, 1, 1, device=dev)
    a = plan.eps(x.to(dev), y.to(dev), noise_level=lvl)
    b = plan.eps(x.to(dev), y.to(dev), noise_level=None, t=37)
    assert rel_err(a, b) < 1e-6


# ----------------------------------------------------------------------------------------------------
# the full loop
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["fp32", "bf16x3", "bf16", "bf16act"])
def test_sampling_cfg1_vs_reference_golden(dev, golden, prec):
    """config_unet.json, cfg-1 clip (2 chunks), full 100 steps, injected noise: per-step eps and final waveform."""
    model, _ = make_model(dev)
    model.noise_estimate_model.precision = prec_id(prec)
    g = golden("sample_cfg1.npz")
    cond = cfg1_condition().to(dev)
    noises = torch.randn(100, 2, 1, L, generator=torch.Generator().manual_seed(1234)).to(dev)
    out, eps_tr, x_tr = model.infer(cond, noises=noises, return_trace=True)
    e100, e50, e1 = (rel_err(eps_tr[100 - t].cpu(), g[f"eps_t{t}"]) for t in (100, 50, 1))
    snr = float(O.sisnr(out.cpu(), g["out"]))
    report(f"sampling cfg1[{prec}]: eps err t=100 {e100:.3e}, t=50 {e50:.3e}, t=1 {e1:.3e}; final SI-SNR {snr:.1f} dB, "
           f"max err {rel_err(out.cpu(), g['out']):.3e}")
    assert e100 <= EPS_BAR[prec] and e50 <= EPS_BAR[prec] and e1 <= EPS_BAR[prec]
    assert snr >= (60.0 if prec in ("fp32", "bf16x3") else SNR_BAR)
    assert float(out.abs().max()) <= 1.0


@pytest.mark.parametrize("variant", ["original", "condition_in", "sr3", "supportive", "conditional"])
def test_sampling_T6_variants_vs_reference_golden(dev, golden, variant):
    from sddm_b200 import PREC_FP32
    model, _ = make_model(dev, T=6, start=1e-4, end=5e-2, variant=variant)
    model.noise_estimate_model.precision = PREC_FP32
    cond = cfg1_condition()[:1].to(dev)
    noises = torch.randn(6, 1, 1, L, generator=torch.Generator().manual_seed(99)).to(dev)
    out = model.infer(cond, noises=noises).cpu()
    e = rel_err(out, golden("sample_T6.npz")[variant])
    report(f"sampling T6 [{variant}] fp32: rel_err={e:.3e}")
    assert e <= 1e-3
    # the step-wise Python loop (one C-ABI call per step) agrees with the fused sampler
    step = model._infer_stepwise(cond, False, noises, 0).cpu()
    assert rel_err(step, out) < 1e-5


def test_batch_invariance_determinism_and_host_api(dev):
    """Rows are independent: a row's result does not depend on its batch, position or the sub-batching; Philox keyed
    by the global row id.  The host-buffer C-ABI call agrees with the device-pointer path."""
    model, _ = make_model(dev, T=5, start=1e-4, end=5e-2)
    net = model.noise_estimate_model
    g = torch.Generator().manual_seed(8)
    cond = (0.1 * torch.randn(7, 1, L, generator=g)).clamp(-1, 1)
    for prec in ("fp32", "bf16x3", "bf16", "bf16act"):
        net.precision = prec_id(prec)
        full = model.infer(cond.to(dev), seed=77)
        again = model.infer(cond.to(dev), seed=77)
        assert torch.equal(full, again), prec
        part = model.infer(cond[3:5].to(dev), seed=77, row0=3)
        assert torch.equal(full[3:5], part), prec
        plan = net.get_plan(model.diffusion)
        host = plan.enhance_host(cond.pin_memory(), "condition_in", seed=77, row0=0, max_rows=3)
        assert torch.equal(host, full.cpu()), prec
        other = model.infer(cond.to(dev), seed=78)
        assert not torch.equal(other, full)
    from sddm_b200.infer import enhance_utterances
    waves = [torch.randn(n, generator=g) * 0.05 for n in (20000, 16448, 300)]
    outs = enhance_utterances(model, waves, batch_chunks=2, seed=5)
    outs2 = enhance_utterances(model, waves, batch_chunks=64, seed=5)
    assert [o.shape[-1] for o in outs] == [20000, 16448, 300]
    assert all(torch.equal(a, b) for a, b in zip(outs, outs2))


def test_back_to_back_forwards(dev):
    """Pipeline stress at the bench size: 400 denoiser forwards of a 64-chunk batch back to back (every persistent CTA runs through
    many ring wraps with the previous kernel's data still in L2).  A ring-parity race of the tcgen05 kernels (DESIGN.md section
    4.2a) showed up only here, once in 30 - 500 forwards, as a launch failure; the result must also stay bit-identical."""
    model, _ = make_model(dev)
    net = model.noise_estimate_model
    net.precision = prec_id("bf16act")
    plan = net.get_plan(model.diffusion)
    g = torch.Generator().manual_seed(31)
    cond = (0.1 * torch.randn(64, 1, L, generator=g)).clamp(-1, 1).to(dev)
    x = torch.randn(64, 1, L, generator=g).to(dev)
    first = plan.eps(cond, x, t=37).clone()
    for it in range(400):
        out = plan.eps(cond, x, t=37)
    torch.cuda.synchronize()
    assert torch.equal(out, first)
    report("stress: 400 back-to-back forwards at B=64 (bf16act): no fault, bit-identical results")


def test_full_size_cfg2(dev, golden):
    """BASELINE cfg 2: 64 chunks, full 100-step schedule, every persistent CTA working through many tiles.
    Pinned to the REFERENCE: rows 0 / 31 / 63 of the batch were run through the reference's SDDM.infer with the same injected
    noise (tests/golden/make_golden_fullsize.py); their per-step eps_hat (t = 100 / 50 / 1) and final waveforms must match
    within the north_star bars in every precision mode.  Size-independent checks on top: the bf16 (tcgen05) and fp32 modes agree
    to >= 40 dB SI-SNR on every row, rows 0-1 reproduce a 2-row run bit for bit, output is clamped to [-1, 1]."""
    model, _ = make_model(dev)
    net = model.noise_estimate_model
    cond, noises = cfg2_inputs()
    cond, noises = cond.to(dev), noises.to(dev)
    gold = golden("fullsize.npz")
    rows = list(CFG2_GOLDEN_ROWS)
    assert gold["cfg2.rows"].tolist() == rows
    outs = {}
    for prec in ("fp32", "bf16x3", "bf16", "bf16act"):
        net.precision = prec_id(prec)
        torch.cuda.synchronize()
        t0 = time.time()
        out, eps_tr, _ = model.infer(cond, noises=noises, return_trace=True)
        torch.cuda.synchronize()
        outs[prec] = out
        report(f"cfg2[{prec}]: 64 chunks x 100 steps in {time.time() - t0:.3f} s (with traces)")
        errs = {t: rel_err(eps_tr[100 - t][rows].cpu(), gold[f"cfg2.eps_t{t}"]) for t in (100, 50, 1)}
        snrs = [float(O.sisnr(out[r:r + 1].cpu(), gold["cfg2.out"][i:i + 1])) for i, r in enumerate(rows)]
        report(f"cfg2[{prec}] rows {rows} vs REFERENCE golden: eps err t=100 {errs[100]:.3e}, t=50 {errs[50]:.3e}, t=1 {errs[1]:.3e}; "
               f"final SI-SNR {min(snrs):.1f} dB (min over rows), max err {rel_err(out[rows].cpu(), gold['cfg2.out']):.3e}")
        assert max(errs.values()) <= EPS_BAR[prec], (prec, errs)
        assert min(snrs) >= (60.0 if prec in ("fp32", "bf16x3") else SNR_BAR), (prec, snrs)
        del eps_tr
        two = model.infer(cond[:2], noises=noises[:, :2].contiguous())
        assert torch.equal(two, outs[prec][:2]), prec
        assert float(outs[prec].abs().max()) <= 1.0 and torch.isfinite(outs[prec]).all()
    for prec in ("bf16", "bf16act"):
        per_row = torch.stack([O.sisnr(outs[prec][i:i + 1].cpu(), outs["fp32"][i:i + 1].cpu()) for i in range(64)])
        report(f"cfg2: {prec} vs fp32 SI-SNR per row: min {float(per_row.min()):.1f} dB, mean {float(per_row.mean()):.1f} dB")
        assert float(per_row.min()) >= SNR_BAR


def test_device_dataset_edge_vs_reference_golden(dev):
    """sddm_chunk_rows / sddm_regroup_rows (the dataset edge on the device) against outputs of the reference's own InferDataset +
    infer_data_collate and the regroup loop of infer.py:81-120 (tests/golden/make_golden_dataset.py): bit-exact, for the whole batch
    and for row ranges (what a rank of a sharded run converts)."""
    import numpy as np
    from conftest import GOLDEN
    from sddm_b200.data_loader import data_loaders as D
    g = np.load(os.path.join(GOLDEN, "dataset.npz"))
    T = int(g["T"])
    order = [str(n).split(".")[0] for n in g["inventory"]]
    waves = [torch.from_numpy(g["wave." + n]).reshape(-1) for n in order]
    noisy = torch.from_numpy(g["noisy"])
    n = noisy.shape[0]
    batch = D.DeviceBatch(waves, T, dev)
    assert batch.n_rows == n
    assert torch.equal(batch.rows(0, n).cpu(), noisy)
    for lo, hi in ((1, 5), (4, 5), (6, n), (0, 1)):
        part = D.DeviceBatch(waves, T, dev, lo, hi)
        assert torch.equal(part.rows(lo, hi).cpu(), noisy[lo:hi]), (lo, hi)
    # regroup: every file, trimmed to its length, equals its input (the reference's reshape(1, -1) of its rows, minus the padding)
    flat = batch.regroup(batch.rows(0, n), 0, n)
    for k, (f, w) in enumerate(zip(batch.split(flat), waves)):
        assert torch.equal(f.cpu().reshape(-1), w)
        ref = torch.from_numpy(g["file%d.signal" % k]).reshape(-1)
        assert torch.equal(f.cpu().reshape(-1), ref[: w.numel()])
    # piecewise regroup (sub-batches) gives the same flat buffer
    flat2 = torch.zeros_like(flat)
    for lo, hi in ((0, 3), (3, 4), (4, n)):
        batch.regroup(batch.rows(lo, hi), lo, hi, flat2)
    assert torch.equal(flat, flat2)
