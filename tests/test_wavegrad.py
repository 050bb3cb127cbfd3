"""cfg 4: WaveGrad denoiser + SDDM_spectrogram sampling loop (fp32 CUDA-core path).

CPU (-m "not gpu"): oracle/wavegrad_oracle.py against goldens of the real reference modules (tests/golden/make_golden_wavegrad.py),
host-mirror surface.  GPU (-m gpu): the CUDA paths through the C ABI (sddm_wg_*).  Tolerance for eps_hat and every block output:
fp32 path 1e-3 of max (measured ~1e-6), tcgen05 bf16 path 2e-2; final sample SI-SNR >= 60 / 40 dB.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import GOLDEN, WAVEGRAD_CASES, cfg4_fullsize_inputs, rel_err, wavegrad_test_module  # noqa: E402
from oracle import sddm_oracle as O  # noqa: E402
from oracle import wavegrad_oracle as WO  # noqa: E402


def report(line):
    """print + append to gpurun_out/parity_report.txt (copied to profiles/ at the end of a round)."""
    print(line)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "parity_report.txt"), "a") as f:
        f.write(line + "\n")


@pytest.fixture(scope="module")
def gold():
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, "wavegrad.npz")).items()}


@pytest.fixture(scope="module")
def sd():
    return {k: v.detach().clone() for k, v in wavegrad_test_module().state_dict().items()}


def si_snr_db(est, ref):
    est, ref = est.double().flatten(), ref.double().flatten()
    s = (est @ ref) / (ref @ ref) * ref
    return float(10 * torch.log10((s @ s) / ((est - s) @ (est - s))))


@pytest.mark.parametrize("tag", list(WAVEGRAD_CASES))
def test_oracle_forward_matches_reference(gold, sd, tag):
    case = WAVEGRAD_CASES[tag]
    trace = {}
    eps = WO.wavegrad_forward(sd, gold[tag + ".spec"], gold[tag + ".audio"], torch.tensor(case["levels"]), trace=trace)
    assert rel_err(eps.reshape(-1), gold[tag + ".eps"].reshape(-1)) < 2e-5
    for k in ("d0", "d2", "d4", "u0", "u2", "u4"):
        assert rel_err(trace[k][:, ::5, ::3], gold[tag + "." + k]) < 2e-5


def test_oracle_sampling_matches_reference(gold, sd):
    x0 = WO.sample_spectrogram(sd, O.make_schedule("linear", 4, 1e-4, 5e-2), gold["sample.spec"], gold["sample.noises"])
    assert si_snr_db(x0, gold["sample.x0"]) > 60.0


def test_mirror_surface():
    from sddm_b200.model import model as M
    from sddm_b200.model.diffusion import GaussianDiffusion
    net = wavegrad_test_module()
    keys = list(net.state_dict().keys())
    assert keys[0] == "downsample.0.weight" and "film.4.output_conv.bias" in keys and keys[-1] == "last_conv.bias"
    assert 15_800_000 < sum(p.numel() for p in net.parameters()) < 16_000_000       # 15.92 M parameters
    d = GaussianDiffusion(schedule="linear", n_timestep=4, linear_start=1e-4, linear_end=5e-2, device="cpu")
    m = M.SDDM_spectrogram(d, net, hop_samples=300)
    assert m.noise_condition == "sqrt_alpha_bar"
    with pytest.raises(RuntimeError):
        m.infer(torch.zeros(1, 128, 3))             # CPU tensors: no fallback


TOL = {"fp32": 1e-3, "bf16": 2e-2}


def _gpu_module(prec):
    from sddm_b200 import _lib
    net = wavegrad_test_module().cuda()
    net.precision = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[prec]
    return net


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", list(WAVEGRAD_CASES))
def test_gpu_eps_vs_reference_golden(built_lib, gold, tag, prec):
    case = WAVEGRAD_CASES[tag]
    net = _gpu_module(prec)
    B, F = case["B"], case["frames"]
    eps = net(gold[tag + ".spec"].cuda(), gold[tag + ".audio"].cuda(), torch.tensor(case["levels"]).cuda()).cpu()
    assert eps.shape == gold[tag + ".eps"].shape                      # torch.squeeze quirk: [T] when B == 1
    plan = net.get_plan()
    errs = {}
    for k in ("d0", "d1", "d2", "d3", "d4", "u0", "u1", "u2", "u3", "u4"):
        errs[k] = rel_err(plan.fetch(k, B, F).cpu()[:, ::5, ::3], gold[tag + "." + k])
    e = rel_err(eps.reshape(-1), gold[tag + ".eps"].reshape(-1))
    report("wavegrad %s %s: eps %.2e  blocks %s" % (tag, prec, e, " ".join("%s %.1e" % kv for kv in errs.items())))
    assert e < TOL[prec] and max(errs.values()) < TOL[prec]


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_gpu_sampling_vs_reference_golden(built_lib, gold, prec):
    from sddm_b200.model import model as M
    from sddm_b200.model.diffusion import GaussianDiffusion
    net = _gpu_module(prec)
    d = GaussianDiffusion(schedule="linear", n_timestep=4, linear_start=1e-4, linear_end=5e-2, device="cuda")
    m = M.SDDM_spectrogram(d, net, hop_samples=300)
    x0 = m.infer(gold["sample.spec"].cuda(), noises=gold["sample.noises"].cuda()).cpu()
    snr = si_snr_db(x0, gold["sample.x0"])
    report("wavegrad sampling %s: SI-SNR vs reference %.1f dB" % (prec, snr))
    assert snr > (60.0 if prec == "fp32" else 40.0)
    a = m.infer(gold["sample.spec"].cuda(), seed=5)
    b = m.infer(gold["sample.spec"].cuda(), seed=5)
    c = m.infer(gold["sample.spec"][1:2].contiguous().cuda(), seed=5, row0=1)
    assert torch.equal(a, b) and torch.equal(a[1:2], c) and float(a.abs().max()) <= 1.0


@pytest.mark.gpu
def test_gpu_ragged_lengths_vs_oracle(built_lib, sd):
    """frame counts whose level lengths are not multiples of the 64-row tile (1 frame: L = 300, 150, 75, 25, 5)."""
    g = torch.Generator().manual_seed(21)
    for prec in ("fp32", "bf16"):
        net = _gpu_module(prec)
        for B, F in ((1, 1), (3, 2), (2, 11)):
            spec, audio = torch.rand(B, 128, F, generator=g), torch.randn(B, 300 * F, generator=g)
            lv = torch.rand(B, generator=g)
            want = WO.wavegrad_forward(sd, spec, audio, lv)
            got = net.get_plan().eps(spec.cuda(), audio.cuda(), noise_level=lv.cuda()).cpu()
            assert rel_err(got, want) < TOL[prec], (prec, B, F)


@pytest.mark.gpu
def test_gpu_full_size_vs_reference_golden_and_row_invariance(built_lib):
    """BASELINE cfg 4 size (spec [128, 107], 32 100 samples).  Rows 0-1 are pinned to the REFERENCE's WaveGrad.forward at this size
    (tests/golden/make_golden_fullsize.py) for both paths; on top: tcgen05 vs fp32 on every row, batch-row invariance, determinism."""
    spec, audio, lv = (t.cuda() for t in cfg4_fullsize_inputs())
    gold = torch.from_numpy(np.load(os.path.join(GOLDEN, "fullsize.npz"))["cfg4.eps_rows01"])
    ref = _gpu_module("fp32").get_plan().eps(spec, audio, noise_level=lv)
    net = _gpu_module("bf16")
    got = net.get_plan().eps(spec, audio, noise_level=lv)
    e32, e16 = rel_err(ref[:2].cpu().reshape(gold.shape), gold), rel_err(got[:2].cpu().reshape(gold.shape), gold)
    e = rel_err(got.cpu(), ref.cpu())
    report("wavegrad full size vs REFERENCE golden: fp32 eps %.2e, tcgen05 eps %.2e; bf16 vs fp32 (3 rows) %.2e" % (e32, e16, e))
    assert e32 < TOL["fp32"] and e16 < TOL["bf16"]
    assert e < 2e-2
    assert torch.equal(got, net.get_plan().eps(spec, audio, noise_level=lv))
    one = net.get_plan().eps(spec[2:3].contiguous(), audio[2:3].contiguous(), noise_level=lv[2:3].contiguous())
    assert torch.equal(one, got[2:3])
