"""Forward diffusion of the training step (forward half): GaussianDiffusion.q_stochastic /
q_stochastic_conditional (reference model/diffusion.py:225-279) and SDDM.forward (model/model.py:29-48).
CPU: oracle vs reference goldens (bit-exact).  GPU: the CUDA kernel through the C ABI vs the goldens (bit-exact: every product and
sum is rounded separately, as the reference's eager ops do), and SDDM.forward vs the oracle's UNet on the same draw."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import GOLDEN, UNET_CFG, rel_err, seed0_state_dict  # noqa: E402
from oracle import sddm_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, "qsample.npz")).items()}


@pytest.fixture(scope="module")
def sch():
    return O.make_schedule("linear", 100, 1e-6, 1e-3)


def test_oracle_matches_reference(gold, sch):
    x_t, level, tt = O.q_stochastic(sch, gold["x0"], gold["noise"], gold["q.t"], gold["q.step"])
    assert torch.equal(x_t, gold["q.x_t"]) and torch.equal(level, gold["q.level"])
    x_t, level, _ = O.q_stochastic(sch, gold["x0"], gold["noise"], gold["qi.t"], None)
    assert torch.equal(x_t, gold["qi.x_t"]) and torch.equal(level, gold["qi.level"])
    x_t, comb, level = O.q_stochastic_conditional(sch, gold["x0"], gold["y"], gold["noise"], gold["qc.t"])
    assert torch.equal(x_t, gold["qc.x_t"]) and torch.equal(comb, gold["qc.combined"]) and torch.equal(level, gold["qc.level"])


@pytest.mark.gpu
def test_gpu_q_sample_bit_exact(built_lib, gold):
    from sddm_b200.model.diffusion import GaussianDiffusion
    d = GaussianDiffusion("linear", 100, 1e-6, 1e-3, device="cuda")
    x0, y, z = gold["x0"].cuda(), gold["y"].cuda(), gold["noise"].cuda()
    x_t, level, tt = d.q_stochastic(x0, z, t=gold["q.t"], random_step=gold["q.step"])
    assert torch.equal(x_t.cpu(), gold["q.x_t"]) and torch.equal(level.cpu(), gold["q.level"])
    assert torch.equal(tt.cpu(), (gold["q.t"] + gold["q.step"]).view(-1, 1, 1))
    x_t, level, _ = d.q_stochastic(x0, z, t_is_integer=True, t=gold["qi.t"])
    assert torch.equal(x_t.cpu(), gold["qi.x_t"]) and torch.equal(level.cpu(), gold["qi.level"])
    x_t, comb, level = d.q_stochastic_conditional(x0, y, z, t=gold["qc.t"])
    assert torch.equal(x_t.cpu(), gold["qc.x_t"]) and torch.equal(comb.cpu(), gold["qc.combined"]) and torch.equal(level.cpu(), gold["qc.level"])
    # Philox draw: N(0,1)-like, deterministic per (seed, row0), returned to the caller
    a = d.q_stochastic(x0, None, seed=3, return_noise=True, t=gold["q.t"], random_step=gold["q.step"])
    b = d.q_stochastic(x0, None, seed=3, return_noise=True, t=gold["q.t"], random_step=gold["q.step"])
    assert torch.equal(a[0], b[0]) and torch.equal(a[3], b[3]) and abs(float(a[3].std()) - 1.0) < 0.05 and abs(float(a[3].mean())) < 0.05
    want = O.q_stochastic(O.make_schedule("linear", 100, 1e-6, 1e-3), gold["x0"], a[3].cpu(), gold["q.t"], gold["q.step"])[0]
    assert torch.equal(a[0].cpu(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("q_transition", ["original", "conditional"])
def test_gpu_sddm_forward_vs_oracle(built_lib, q_transition):
    """SDDM.forward = forward-diffusion draw + eps_hat: (predicted, noise) as the trainer's loss consumes them."""
    from sddm_b200 import PREC_FP32
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM
    sd, net = seed0_state_dict()
    net.precision = PREC_FP32
    d = GaussianDiffusion("linear", 100, 1e-6, 1e-3, device="cuda")
    m = SDDM(d, net, q_transition=q_transition, p_transition="conditional" if q_transition == "conditional" else "original").cuda().eval()
    g = torch.Generator().manual_seed(2)
    target = (0.1 * torch.randn(2, 1, 16448, generator=g)).clamp(-1, 1)
    cond = (target + 0.03 * torch.randn(2, 1, 16448, generator=g)).clamp(-1, 1)
    torch.manual_seed(11)
    predicted, noise = m(target.cuda(), cond.cuda(), seed=5)
    assert predicted.shape == target.shape and noise.shape == target.shape
    # re-derive the draw on the CPU: same torch RNG stream for t / step, the Philox noise comes back from the call
    sch = O.make_schedule("linear", 100, 1e-6, 1e-3)
    torch.manual_seed(11)
    if q_transition == "original":
        t = torch.randint(1, 101, [2], device="cuda").cpu()
        step = torch.rand(2, device="cuda").cpu()
        x_t, level, _ = O.q_stochastic(sch, target, noise.cpu(), t, step)
        want_noise = noise.cpu()
    else:
        t = torch.randint(1, 101, (2, 1, 1), device="cuda").cpu()
        # the Philox draw itself is not returned by the conditional branch: recover it from combined_noise is ill-posed, so
        # check the eps_hat against the oracle on the x_t the library produced
        x_t = d.q_stochastic_conditional(target.cuda(), cond.cuda(), None, t=t, seed=5)[0].cpu()
        level = sch["sqrt_alpha_bar"][t]
        want_noise = None
    want = O.unet_forward(sd, dict(UNET_CFG), cond, x_t, level.view(2, 1, 1))
    assert rel_err(predicted.cpu(), want) < 1e-3
    if want_noise is not None:
        assert torch.equal(noise.cpu(), want_noise)
