"""Golden vectors at BASELINE.json's FULL sizes, produced by the REAL reference modules (build container only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_fullsize.py

The small-size goldens exercise at most 2 tiles per CTA of the persistent kernels; these pin the sizes the benchmark runs:

  * cfg 2 ... rows 0, 31 and 63 of the 64-chunk batch (cond seed 0, injected noise seed 1234) through the reference's
              SDDM.infer, all 100 steps (rows are independent: GroupNorm is per sample)        model/model.py:105-124
              -> final waveform + eps_hat at t = 100 / 50 / 1 per row
  * cfg 5 ... one eps_hat of the config-shaped DiffWave on one 10 s utterance (spec [513, 626], 160 256 samples)
                                                                                               model/diffwave.py:133-155
  * cfg 4 ... one eps_hat of WaveGrad at F = 107 frames (32 100 samples), 2 utterances          model/wavegrad.py:167-179

Inputs are regenerated from seeds by the tests (torch CPU generators are deterministic); only outputs are stored.
"""
import importlib.util
import os
import sys
import time

import numpy as np
import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import model.diffusion as ref_diffusion  # noqa: E402
import model.model as ref_model          # noqa: E402
import model.network as ref_network      # noqa: E402
from model.diffwave import DiffWave as RefDiffWave  # noqa: E402
from model.wavegrad import WaveGrad as RefWaveGrad  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(1, ROOT)
_spec = importlib.util.spec_from_file_location("sddm_conftest", os.path.join(ROOT, "tests", "conftest.py"))
_conftest = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_conftest)
OUT = os.path.dirname(os.path.abspath(__file__))
L = 16448
CFG2_ROWS = _conftest.CFG2_GOLDEN_ROWS


def cfg2():
    torch.manual_seed(0)
    diff = ref_diffusion.GaussianDiffusion(schedule="linear", n_timestep=100, linear_start=1e-6, linear_end=1e-3, device="cpu")
    net = ref_network.UNetModified2(num_samples=L, in_channel=2, out_channel=1, inner_channel=32, norm_groups=32,
                                    channel_mults=[1, 2, 3, 4, 5], res_blocks=1, dropout=0, segment_len=128, segment_stride=64)
    model = ref_model.SDDM(diff, net, p_transition="condition_in").eval()
    cond, noises = _conftest.cfg2_inputs()
    rows = list(CFG2_ROWS)
    c, z = cond[rows].contiguous(), noises[:, rows].contiguous()
    trace = []
    orig_fwd = net.forward

    def traced(xc, yt, lvl):
        e = orig_fwd(xc, yt, lvl)
        trace.append(e.detach().clone())
        return e

    net.forward = traced
    k = [0]

    def inj(like, **kw):
        v = z[k[0]].reshape(like.shape)
        k[0] += 1
        return v

    orig = torch.randn_like
    torch.randn_like = inj
    t0 = time.time()
    try:
        with torch.no_grad():
            out = model.infer(c)
    finally:
        torch.randn_like = orig
        net.forward = orig_fwd
    assert k[0] == 100 and len(trace) == 100
    print("cfg2 rows", rows, "reference SDDM.infer: %.1f s" % (time.time() - t0), float(out.abs().max()))
    return {"cfg2.rows": np.asarray(rows), "cfg2.out": out.numpy(), "cfg2.eps_t100": trace[0].numpy(), "cfg2.eps_t50": trace[50].numpy(),
            "cfg2.eps_t1": trace[99].numpy()}


def cfg5():
    case = _conftest.DIFFWAVE_CASES["full"]
    torch.manual_seed(0)
    net = RefDiffWave(num_samples=-1, num_timesteps=200, freq_bins=513, residual_channels=64, residual_layers=30, dilation_cycle_length=10)
    with torch.no_grad():
        net.output_projection.weight.copy_(0.1 * torch.randn(net.output_projection.weight.shape, generator=torch.Generator().manual_seed(1)))
    net.eval()
    mirror = _conftest.diffwave_test_module(case)
    for k_, v in net.state_dict().items():
        assert torch.equal(v, mirror.state_dict()[k_]), k_
    spec, audio, step = _conftest.cfg5_fullsize_inputs()
    t0 = time.time()
    with torch.no_grad():
        eps = net(spec[:1], audio[:1], step[:1])
    print("cfg5 full-size reference eps_hat: %.1f s" % (time.time() - t0), tuple(eps.shape), float(eps.std()))
    return {"cfg5.eps_row0": eps.numpy()}


def cfg4():
    torch.manual_seed(0)
    net = RefWaveGrad()
    mirror = _conftest.wavegrad_test_module()
    net.load_state_dict(mirror.state_dict())
    net.eval()
    spec, audio, lv = _conftest.cfg4_fullsize_inputs()
    t0 = time.time()
    with torch.no_grad():
        eps = net(spec[:2], audio[:2], lv[:2])
    print("cfg4 full-size reference eps_hat: %.1f s" % (time.time() - t0), tuple(eps.shape), float(eps.std()))
    return {"cfg4.eps_rows01": eps.numpy()}


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    out = {}
    out.update(cfg4())
    out.update(cfg5())
    out.update(cfg2())
    path = os.path.join(OUT, "fullsize.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
