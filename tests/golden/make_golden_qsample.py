"""Golden vectors for the forward-diffusion draws (training step), from the REAL reference (build container only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_qsample.py
Pinned: GaussianDiffusion.q_stochastic / q_stochastic_conditional, model/diffusion.py:225-279.  The reference draws t (and the
uniform step) with torch.randint / torch.rand; they are re-drawn here under the same seed and stored with the outputs."""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import model.diffusion as ref_diffusion  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    d = ref_diffusion.GaussianDiffusion(schedule="linear", n_timestep=100, linear_start=1e-6, linear_end=1e-3, device="cpu")
    g = torch.Generator().manual_seed(31)
    B, L = 5, 4096
    x0 = (0.2 * torch.randn(B, 1, L, generator=g)).clamp(-1, 1)
    y = (x0 + 0.05 * torch.randn(B, 1, L, generator=g)).clamp(-1, 1)
    noise = torch.randn(B, 1, L, generator=g)
    out = {"x0": x0.numpy(), "y": y.numpy(), "noise": noise.numpy()}
    # q_stochastic, continuous noise level
    torch.manual_seed(7)
    x_t, level, tt = d.q_stochastic(x0, noise)
    torch.manual_seed(7)
    t = torch.randint(1, 101, [B])
    step = torch.rand(B)
    assert torch.equal((t + step).view(B, 1, 1), tt)
    out.update({"q.t": t.numpy(), "q.step": step.numpy(), "q.x_t": x_t.numpy(), "q.level": level.numpy()})
    # q_stochastic, integer steps
    torch.manual_seed(8)
    x_t, level, tt = d.q_stochastic(x0, noise, t_is_integer=True)
    torch.manual_seed(8)
    t = torch.randint(1, 101, [B])
    assert torch.equal(t.view(B, 1, 1), tt)
    out.update({"qi.t": t.numpy(), "qi.x_t": x_t.numpy(), "qi.level": level.numpy()})
    # conditional
    torch.manual_seed(9)
    x_t, comb, level = d.q_stochastic_conditional(x0, y, noise)
    torch.manual_seed(9)
    t = torch.randint(1, 101, (B, 1, 1))
    out.update({"qc.t": t.numpy(), "qc.x_t": x_t.numpy(), "qc.combined": comb.numpy(), "qc.level": level.numpy()})
    np.savez_compressed(os.path.join(OUT, "qsample.npz"), **out)
    print("wrote qsample.npz", os.path.getsize(os.path.join(OUT, "qsample.npz")))


if __name__ == "__main__":
    main()
