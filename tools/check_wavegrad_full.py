#!/usr/bin/env python
"""One-off check at the bench size (8 utterances, spec [128,107]): WaveGrad tcgen05 path vs the fp32 path, row by row."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from conftest import rel_err, wavegrad_test_module  # noqa: E402
from sddm_b200 import _lib  # noqa: E402

B, F = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 107
g = torch.Generator().manual_seed(3)
spec, audio = torch.rand(B, 128, F, generator=g).cuda(), torch.randn(B, 300 * F, generator=g).cuda()
lv = torch.rand(B, generator=g).cuda()
outs = {}
for prec, code in (("fp32", _lib.PREC_FP32), ("bf16", _lib.PREC_BF16)):
    net = wavegrad_test_module().cuda()
    net.precision = code
    outs[prec] = net.get_plan().eps(spec, audio, noise_level=lv).cpu()
    if prec == "bf16":
        again = net.get_plan().eps(spec, audio, noise_level=lv).cpu()
        print("deterministic:", torch.equal(again, outs[prec]))
print("B=%d: bf16 vs fp32 overall %.2e, per row %s" % (B, rel_err(outs["bf16"], outs["fp32"]),
                                                       " ".join("%.1e" % rel_err(outs["bf16"][i], outs["fp32"][i]) for i in range(B))))
