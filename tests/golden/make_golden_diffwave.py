"""Golden vectors for the cfg-5 path, produced by the REAL reference modules (build container only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_diffwave.py

Pinned (reference file:line):
  * default-init weights ... DiffWave.__init__ under torch.manual_seed(0)   model/diffwave.py:111-131 (must equal the host mirror's)
  * eps_hat ................ DiffWave.forward                                model/diffwave.py:133-155
  * full sampling .......... SDDM_spectrogram.infer                          model/model.py:206-257 (torch.randn / randn_like injected)
The reference zero-initialises output_projection.weight (diffwave.py:131), which would make eps_hat a constant; every case
below replaces it by 0.1 * N(0,1) (generator seed 1) — tests/conftest.py: diffwave_test_module does the same.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import model.diffusion as ref_diffusion  # noqa: E402
import model.model as ref_model          # noqa: E402
from model.diffwave import DiffWave as RefDiffWave  # noqa: E402

import importlib.util  # noqa: E402
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(1, ROOT)
_spec = importlib.util.spec_from_file_location("sddm_conftest", os.path.join(ROOT, "tests", "conftest.py"))
_conftest = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_conftest)
DIFFWAVE_CASES, diffwave_test_module = _conftest.DIFFWAVE_CASES, _conftest.diffwave_test_module

OUT = os.path.dirname(os.path.abspath(__file__))


def ref_module(case):
    torch.manual_seed(0)
    net = RefDiffWave(num_samples=-1, num_timesteps=200, freq_bins=case["freq_bins"], residual_channels=64,
                      residual_layers=case["residual_layers"], dilation_cycle_length=case["dilation_cycle_length"])
    with torch.no_grad():
        net.output_projection.weight.copy_(0.1 * torch.randn(net.output_projection.weight.shape, generator=torch.Generator().manual_seed(1)))
    return net.eval()


def main():
    torch.set_num_threads(8)
    out = {}
    for tag, case in DIFFWAVE_CASES.items():
        net = ref_module(case)
        mirror = diffwave_test_module(case)
        sd_r, sd_m = net.state_dict(), mirror.state_dict()
        assert list(sd_r.keys()) == list(sd_m.keys()), "state_dict keys differ"
        for k in sd_r:
            assert torch.equal(sd_r[k], sd_m[k]), "mirror init differs from the reference at " + k
        assert torch.equal(net.diffusion_embedding.embedding_vector, mirror.diffusion_embedding.embedding_vector)
        g = torch.Generator().manual_seed(case["seed"])
        B, frames, F = case["B"], case["frames"], case["freq_bins"]
        spec = torch.rand(B, F, frames, generator=g) * 0.7
        audio = torch.randn(B, 1, 256 * frames, generator=g)
        step = torch.tensor(case["steps"], dtype=torch.float32).reshape(B, 1, 1)
        grabbed = {}
        hooks = [net.spectrogram_upsampler.register_forward_hook(lambda m, i, o: grabbed.__setitem__("up", o))]
        for i in case["probe_layers"]:
            hooks.append(net.residual_layers[i].register_forward_hook(lambda m, inp, o, i=i: grabbed.__setitem__("x%d" % i, o[0])))
        with torch.no_grad():
            eps = net(spec, audio, step)
        for h in hooks:
            h.remove()
        out[tag + ".spec"] = spec.numpy()
        out[tag + ".audio"] = audio.numpy()
        out[tag + ".eps"] = eps.numpy()
        out[tag + ".up_last"] = grabbed["up"][B - 1, :, ::29].numpy()                 # [F, T/29] of the last utterance
        for i in case["probe_layers"]:
            out[tag + ".x%d" % i] = grabbed["x%d" % i][:, ::4, ::3].numpy()          # [B, 16, T/3]
        print(tag, "eps", float(eps.abs().max()), float(eps.std()))

    # full sampling: 6 steps, time_step conditioning, config-shaped network
    case = DIFFWAVE_CASES["full"]
    net = ref_module(case)
    Tn, B, frames = 6, 2, 4
    d = ref_diffusion.GaussianDiffusion(schedule="linear", n_timestep=Tn, linear_start=1e-4, linear_end=5e-2, device="cpu")
    g = torch.Generator().manual_seed(77)
    spec = torch.rand(B, case["freq_bins"], frames, generator=g) * 0.7
    noises = torch.randn(Tn, B, 1, 256 * frames, generator=g)
    for cond_kind in ("time_step", "sqrt_alpha_bar"):
        model = ref_model.SDDM_spectrogram(d, net, hop_samples=256, noise_condition=cond_kind).eval()
        k = [0]

        def next_noise(*a, **kw):
            z = noises[k[0]]
            k[0] += 1
            return z.clone()
        orig = torch.randn, torch.randn_like
        torch.randn = next_noise
        torch.randn_like = lambda like, **kw: next_noise().reshape(like.shape)
        try:
            with torch.no_grad():
                x0 = model.infer(spec)
        finally:
            torch.randn, torch.randn_like = orig
        assert k[0] == Tn, k
        out["sample.%s.x0" % cond_kind] = x0.numpy()
        print("sample", cond_kind, float(x0.abs().max()), float(x0.std()))
    out["sample.spec"] = spec.numpy()
    out["sample.noises"] = noises.numpy()
    np.savez_compressed(os.path.join(OUT, "diffwave.npz"), **out)
    print("wrote", os.path.join(OUT, "diffwave.npz"), os.path.getsize(os.path.join(OUT, "diffwave.npz")))


if __name__ == "__main__":
    main()
