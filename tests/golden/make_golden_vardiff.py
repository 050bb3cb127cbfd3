"""Golden vectors of the SNR-adaptive diffusion (schedule / start / posterior half): outputs of the
reference's VariableGaussianDiffusion itself (model/diffusion.py:329-446), imported read-only from /root/reference on the CPU.
Build container only:
    python tests/golden/make_golden_vardiff.py
Writes tests/golden/vardiff.npz: per-frame SNR estimates, the schedules they induce, get_x_T / get_noise_level / p_transition /
q_stochastic outputs with the injected noise (torch.randn_like / torch.randint are patched to return recorded draws)."""
import os
import sys
from unittest import mock

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from model.diffusion import VariableGaussianDiffusion  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
g = torch.Generator().manual_seed(404)
B, N, L, T = 3, 7, 16, 100
diff = VariableGaussianDiffusion(n_timestep=T, snr_estimate_scale=100, device="cpu")
snr = torch.empty(B, N).uniform_(-8.0, 30.0, generator=g)          # dB, per frame
cond = (0.3 * torch.randn(B, 1, N, L, generator=g)).clamp(-1, 1)
out = {"snr": snr.numpy(), "cond": cond.numpy(), "T": np.asarray(T), "scale": np.asarray(100.0)}
betas, alpha_bar = diff.get_beta_schedule(snr)
out["betas"] = betas.numpy()            # [B, 1, N, T + 1]
out["alpha_bar"] = alpha_bar.numpy()
z0 = torch.randn(B, 1, N, L, generator=g)
with mock.patch.object(torch, "randn_like", lambda *a, **k: z0.clone()):
    out["x_T"] = diff.get_x_T(cond, snr).numpy()
out["z_x_T"] = z0.numpy()
for t in (100, 50, 2, 1):
    out["noise_level_t%d" % t] = diff.get_noise_level(t, snr).numpy()
    x_t = (0.5 * torch.randn(B, 1, N, L, generator=g)).clamp(-1, 1)
    eps = torch.randn(B, 1, N, L, generator=g)
    z = torch.randn(B, 1, N, L, generator=g)
    with mock.patch.object(torch, "randn_like", lambda *a, **k: z.clone()):
        y = diff.p_transition(x_t.clone(), t, snr, eps)
    out["post_t%d.x_t" % t] = x_t.numpy()
    out["post_t%d.eps" % t] = eps.numpy()
    out["post_t%d.z" % t] = z.numpy()
    out["post_t%d.out" % t] = y.numpy()
for t in (100, 37, 1):
    x0 = (0.3 * torch.randn(B, 1, N, L, generator=g)).clamp(-1, 1)
    noise = torch.randn(B, 1, N, L, generator=g)
    with mock.patch.object(torch, "randint", lambda *a, **k: torch.tensor([t])):
        x_t, nl, tt = diff.q_stochastic(x0, noise, snr)
    assert int(tt) == t
    out["q_t%d.x0" % t] = x0.numpy()
    out["q_t%d.noise" % t] = noise.numpy()
    out["q_t%d.x_t" % t] = x_t.numpy()
    out["q_t%d.noise_level" % t] = nl.numpy()
np.savez_compressed(os.path.join(OUT, "vardiff.npz"), **out)
print({k: v.shape for k, v in out.items()})
print(os.path.getsize(os.path.join(OUT, "vardiff.npz")))
