"""SNR-adaptive diffusion (diffusion half): oracle vs the reference goldens on the CPU; the CUDA kernels
(csrc/var_diffusion.cu, through the C ABI) vs the same goldens on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import var_diffusion_oracle as VO  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "vardiff.npz")


def report(line):
    """print + append to gpurun_out/parity_report.txt (copied to profiles/ at the end of a round)."""
    print(line)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.txt"), "a") as f:
        f.write(line + "\n")


@pytest.fixture(scope="module")
def gold():
    g = np.load(GOLD)
    return {k: torch.from_numpy(g[k]) if g[k].ndim else g[k] for k in g.files}


def test_oracle_schedule_bit_exact(gold):
    T, scale = int(gold["T"]), float(gold["scale"])
    betas, ab = VO.beta_schedule(gold["snr"], T, scale)
    assert torch.equal(betas, gold["betas"])
    assert torch.equal(ab, gold["alpha_bar"])


def test_oracle_steps_bit_exact(gold):
    T, scale = int(gold["T"]), float(gold["scale"])
    snr = gold["snr"]
    assert torch.equal(VO.x_T(gold["cond"], snr, gold["z_x_T"], T, scale), gold["x_T"])
    for t in (100, 50, 2, 1):
        assert torch.equal(VO.noise_level(snr, t, T, scale), gold["noise_level_t%d" % t])
        y = VO.p_transition(gold["post_t%d.x_t" % t], t, snr, gold["post_t%d.eps" % t], gold["post_t%d.z" % t], T, scale)
        assert torch.equal(y, gold["post_t%d.out" % t]), t
    for t in (100, 37, 1):
        x_t, s = VO.q_sample(gold["q_t%d.x0" % t], gold["q_t%d.noise" % t], snr, t, T, scale)
        assert torch.equal(x_t, gold["q_t%d.x_t" % t]) and torch.equal(s, gold["q_t%d.noise_level" % t])


def test_host_mirror_has_no_cpu_fallback(gold):
    """The product path runs on CUDA tensors only: CPU tensors raise, shapes are checked before anything is launched."""
    from sddm_b200.model.diffusion import VariableGaussianDiffusion
    d = VariableGaussianDiffusion(n_timestep=100, snr_estimate_scale=100, device="cpu")
    snr, cond = gold["snr"], gold["cond"]
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        d.get_beta_schedule(snr)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        d.get_x_T(cond, snr)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        d.p_transition(cond, 5, snr, cond)
    with pytest.raises(NotImplementedError):
        d.q_stochastic(cond, cond, snr, t_is_integer=False)


# ----------------------------------------------------------------------------------------------------
# CUDA kernels through the C ABI (host mirror VariableGaussianDiffusion)
# ----------------------------------------------------------------------------------------------------
TOL = 2e-6     # fp32 path; every output is in [-1, 1] / a schedule term in [0, 1].  The only operation that is not reproduced bit for
               # bit is 10^x: the reference's vectorised CPU powf is allowed 1 ulp, the kernel rounds the fp64 result correctly


@pytest.fixture(scope="module")
def vdiff():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import __graft_entry__ as ge
    ge.build()
    from sddm_b200.model.diffusion import VariableGaussianDiffusion
    return VariableGaussianDiffusion(n_timestep=100, snr_estimate_scale=100, device="cuda")


@pytest.mark.gpu
def test_gpu_schedule_and_steps_vs_reference_golden(vdiff, gold):
    dev = torch.device("cuda:0")
    snr = gold["snr"].to(dev)
    betas, ab = vdiff.get_beta_schedule(snr)
    eb = float((betas.cpu() - gold["betas"]).abs().max()), float((ab.cpu() - gold["alpha_bar"]).abs().max())
    exact = float((betas.cpu() == gold["betas"]).float().mean()), float((ab.cpu() == gold["alpha_bar"]).float().mean())
    report("vardiff schedule: max abs err betas %.2e alpha_bar %.2e; bit-identical %.3f / %.3f of the entries" % (eb + exact))
    assert max(eb) <= TOL
    x_T = vdiff.get_x_T(gold["cond"].to(dev), snr, noise=gold["z_x_T"].to(dev)).cpu()
    assert float((x_T - gold["x_T"]).abs().max()) <= TOL
    worst = 0.0
    for t in (100, 50, 2, 1):
        nl = vdiff.get_noise_level(t, snr).cpu()
        assert nl.shape == gold["noise_level_t%d" % t].shape
        assert float((nl - gold["noise_level_t%d" % t]).abs().max()) <= TOL
        y = vdiff.p_transition(gold["post_t%d.x_t" % t].to(dev), t, snr, gold["post_t%d.eps" % t].to(dev), noise=gold["post_t%d.z" % t].to(dev)).cpu()
        e = float((y - gold["post_t%d.out" % t]).abs().max())
        worst = max(worst, e)
        assert e <= TOL, (t, e)
        assert float(y.abs().max()) <= 1.0
    for t in (100, 37, 1):
        x_t, s, tt = vdiff.q_stochastic(gold["q_t%d.x0" % t].to(dev), gold["q_t%d.noise" % t].to(dev), snr, t=t)
        assert int(tt) == t
        assert float((x_t.cpu() - gold["q_t%d.x_t" % t]).abs().max()) <= TOL
        assert float((s.cpu() - gold["q_t%d.noise_level" % t]).abs().max()) <= TOL
    report("vardiff p_transition: max abs err %.2e over t in (100, 50, 2, 1)" % worst)


@pytest.mark.gpu
def test_gpu_reverse_loop_properties(vdiff):
    """Size-independent properties at a BASELINE-sized frame grid (64 rows x 256 frames x 128 samples): frames with a high SNR
    estimate (+80 dB) get far less noise than frames below 0 dB, a frame's result depends only on its own SNR (rows permuted -> results permuted), the Philox stream is
    keyed by the global frame (a sub-batch reproduces its rows), the loop stays in [-1, 1] and is deterministic."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(9)
    B, N, L = 64, 256, 128
    cond = (0.2 * torch.randn(B, 1, N, L, generator=g)).clamp(-1, 1).to(dev)
    snr = torch.empty(B, N).uniform_(-5.0, 25.0, generator=g)
    snr[0, :] = 80.0
    snr = snr.to(dev)

    def run(c, s, row0=0):
        x = vdiff.get_x_T(c, s, seed=5, row0=row0)
        for t in range(vdiff.num_timesteps, 0, -1):
            x = vdiff.p_transition(x, t, s, torch.zeros_like(x), seed=5, row0=row0)     # eps_hat = 0: the schedule alone
        return x

    out = run(cond, snr)
    assert torch.isfinite(out).all() and float(out.abs().max()) <= 1.0
    assert torch.equal(out, run(cond, snr))
    dev_rms = (out - cond).square().mean(dim=-1).sqrt().reshape(B, N)                    # per frame: how far the loop moved it
    noisy = snr < 0.0
    report("vardiff loop (eps_hat = 0): rms move of +80 dB frames %.2e, of < 0 dB frames %.2e" % (float(dev_rms[0].mean()), float(dev_rms[noisy].mean())))
    assert float(dev_rms[0].max()) < 2e-2 and float(dev_rms[noisy].mean()) > 3.0 * float(dev_rms[0].mean())   # the schedule follows the SNR
    assert torch.equal(run(cond[10:13].contiguous(), snr[10:13].contiguous(), row0=10), out[10:13])
    perm = torch.randperm(B, generator=g).to(dev)
    nl = vdiff.get_noise_level(37, snr)
    assert torch.equal(vdiff.get_noise_level(37, snr[perm].contiguous()), nl[perm])
