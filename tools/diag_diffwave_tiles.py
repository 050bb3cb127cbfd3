#!/usr/bin/env python
"""Debug aid: which 128-sample tiles of the residual stream differ between the tcgen05 and the fp32 path for a short DiffWave."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from sddm_b200 import _lib  # noqa: E402
from sddm_b200.model.network import DiffWave  # noqa: E402

layers, B, frames = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
T = 256 * frames


def mod(prec):
    torch.manual_seed(0)
    n = DiffWave(freq_bins=513, residual_layers=layers, dilation_cycle_length=10).cuda()
    with torch.no_grad():
        n.output_projection.weight.copy_(0.1 * torch.randn(n.output_projection.weight.shape, generator=torch.Generator().manual_seed(1)))
    n.precision = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[prec]
    return n


g = torch.Generator().manual_seed(12)
spec = (torch.rand(B, 513, frames, generator=g) * 0.7).cuda()
audio = torch.randn(B, 1, T, generator=g).cuda()
step = torch.full((B, 1, 1), 77.0).cuda()
out = {}
for prec in ("fp32", "bf16"):
    net = mod(prec)
    eps = net(spec, audio, step)
    out[prec] = (net.get_plan().fetch("x", B, frames).reshape(B * T // 128, 128 * 64).cpu(), eps.cpu(),
                 net.get_plan().fetch("z%d" % (layers - 1) if prec == "bf16" else "z", B, frames).reshape(B * T // 128, 128 * 64).cpu())
xf, xb = out["fp32"][0], out["bf16"][0]
err = (xf - xb).abs().amax(dim=1) / xf.abs().max()
bad = (err > 0.05).nonzero().flatten()
print("layers=%d B=%d frames=%d tiles=%d: bad tiles %d, max tile err %.2e, eps err %.2e" %
      (layers, B, frames, xf.shape[0], bad.numel(), float(err.max()), float((out["fp32"][1] - out["bf16"][1]).abs().max() / out["fp32"][1].abs().max())))
if bad.numel():
    b = bad.tolist()
    zf, zb = out["fp32"][2], out["bf16"][2]
    for tl in b[:3]:
        ex = (xf[tl] - xb[tl]).abs().reshape(128, 64)
        ez = (zf[tl] - zb[tl]).abs().reshape(128, 64)
        print("tile %d: x err rows %s cols %s | z err max %.2e rows %s cols %s" %
              (tl, (ex.amax(1) > 0.05).nonzero().flatten().tolist()[:40], (ex.amax(0) > 0.05).nonzero().flatten().tolist()[:70], float(ez.max()),
               (ez.amax(1) > 0.05).nonzero().flatten().tolist()[:40], (ez.amax(0) > 0.05).nonzero().flatten().tolist()[:70]))
    print("first bad", b[:24])
    print("bad mod 148:", sorted(set(x % 148 for x in b))[:40])
    print("bad // 148:", sorted(set(x // 148 for x in b)))
