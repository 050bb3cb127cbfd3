#!/usr/bin/env python
"""Debug aid: row 0 of a DiffWave eps_hat evaluation must not depend on the batch it is part of.  Compares the cached conditioner,
the residual stream and eps_hat of row 0 between B = 1 and B = 2 (tcgen05 path) and against the fp32 path."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from conftest import DIFFWAVE_CASES, diffwave_test_module, rel_err  # noqa: E402
from sddm_b200 import _lib  # noqa: E402

case = DIFFWAVE_CASES["full"]


def mod(prec):
    n = diffwave_test_module(case).cuda()
    n.precision = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[prec]
    return n


g = torch.Generator().manual_seed(12)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 300
T = 256 * frames
spec = (torch.rand(2, 513, frames, generator=g) * 0.7).cuda()
audio = torch.randn(2, 1, T, generator=g).cuda()
step = torch.tensor([150.0, 20.0]).reshape(2, 1, 1).cuda()
res = {}
for B in (1, 2):
    for prec in ("bf16", "fp32"):
        net = mod(prec)
        eps = net(spec[:B].contiguous(), audio[:B].contiguous(), step[:B].contiguous())
        plan = net.get_plan()
        d = {"eps": eps[0].cpu()}
        for name, w in (("cond0", 128), ("cond29", 128), ("x", 64)):
            d[name] = plan.fetch(name, B, frames).reshape(B, T, w)[0].cpu()
        res[(B, prec)] = d
for name in ("cond0", "cond29", "x", "eps"):
    a, b = res[(1, "bf16")][name], res[(2, "bf16")][name]
    f = res[(1, "fp32")][name]
    bad = (a != b).reshape(a.shape[0] if a.dim() > 1 else -1, -1).any(dim=-1) if a.dim() > 1 else (a != b).flatten()
    idx = bad.nonzero().flatten()
    print("%-7s B1 vs B2 (bf16, row 0): %.2e   B1 bf16 vs fp32: %.2e   B2 bf16 vs fp32: %.2e   fp32 B1 vs B2: %.2e   differing rows %d (first %s, last %s)"
          % (name, rel_err(a, b), rel_err(a, f), rel_err(b, f), rel_err(res[(1, "fp32")][name], res[(2, "fp32")][name]), idx.numel(),
             idx[:4].tolist(), idx[-4:].tolist()), flush=True)
