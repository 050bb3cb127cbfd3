"""Golden vectors for the cfg-4 path, produced by the REAL reference modules (build container only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_wavegrad.py

Pinned (reference file:line):
  * default-init weights ... WaveGrad.__init__ under torch.manual_seed(0)   model/wavegrad.py:140-165 (must equal the host mirror's)
  * eps_hat ................ WaveGrad.forward                                model/wavegrad.py:167-179
  * full sampling .......... SDDM_spectrogram.infer                          model/model.py:206-257, with the squeeze adapter the
                             shipped wrapper lacks (it passes [B,1,T] where WaveGrad.forward needs [B,T])
The reference zero-initialises every bias; all cases replace them by 0.05 * N(0,1) (tests/conftest.py: wavegrad_test_module).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import model.diffusion as ref_diffusion  # noqa: E402
import model.model as ref_model          # noqa: E402
from model.wavegrad import WaveGrad as RefWaveGrad  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(1, ROOT)
_spec = importlib.util.spec_from_file_location("sddm_conftest", os.path.join(ROOT, "tests", "conftest.py"))
_conftest = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_conftest)
OUT = os.path.dirname(os.path.abspath(__file__))


class Squeezed(torch.nn.Module):
    """The adapter the shipped SDDM_spectrogram lacks: [B,1,T] in / out around WaveGrad.forward([B,T])."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, spectrogram, x_t, noise_level):
        return self.net(spectrogram, x_t.squeeze(1), noise_level).reshape(x_t.shape)


def main():
    torch.set_num_threads(8)
    torch.manual_seed(0)
    net = RefWaveGrad()
    mirror = _conftest.wavegrad_test_module()
    torch.manual_seed(0)
    fresh = RefWaveGrad()
    sd_r, sd_m = fresh.state_dict(), mirror.state_dict()
    assert list(sd_r.keys()) == list(sd_m.keys()), "state_dict keys differ"
    for k in sd_r:
        if not k.endswith(".bias"):
            assert torch.equal(sd_r[k], sd_m[k]), "mirror init differs from the reference at " + k
    net.load_state_dict(sd_m)
    net.eval()
    out = {}
    for tag, case in _conftest.WAVEGRAD_CASES.items():
        g = torch.Generator().manual_seed(case["seed"])
        B, F = case["B"], case["frames"]
        spec = torch.rand(B, 128, F, generator=g)
        audio = torch.randn(B, 300 * F, generator=g)
        nl = torch.tensor(case["levels"], dtype=torch.float32)
        grabbed = {}
        hooks = [net.downsample[i].register_forward_hook(lambda m, a, o, i=i: grabbed.__setitem__("d%d" % i, o)) for i in range(5)]
        hooks += [net.upsample[i].register_forward_hook(lambda m, a, o, i=i: grabbed.__setitem__("u%d" % i, o)) for i in range(5)]
        with torch.no_grad():
            eps = net(spec, audio, nl)
        for h in hooks:
            h.remove()
        out[tag + ".spec"], out[tag + ".audio"], out[tag + ".eps"] = spec.numpy(), audio.numpy(), eps.numpy()
        for k, v in grabbed.items():
            out[tag + "." + k] = v[:, ::5, ::3].numpy()
        print(tag, tuple(eps.shape), float(eps.abs().max()), float(eps.std()))
    Tn, B, F = 4, 2, 3
    d = ref_diffusion.GaussianDiffusion(schedule="linear", n_timestep=Tn, linear_start=1e-4, linear_end=5e-2, device="cpu")
    g = torch.Generator().manual_seed(78)
    spec = torch.rand(B, 128, F, generator=g)
    noises = torch.randn(Tn, B, 1, 300 * F, generator=g)
    model = ref_model.SDDM_spectrogram(d, Squeezed(net), hop_samples=300).eval()
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
    assert k[0] == Tn
    out["sample.spec"], out["sample.noises"], out["sample.x0"] = spec.numpy(), noises.numpy(), x0.numpy()
    print("sample", float(x0.abs().max()), float(x0.std()))
    path = os.path.join(OUT, "wavegrad.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
