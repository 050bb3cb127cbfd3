"""Diagnostic: one full-size cfg-2 sampling run (64 chunks x 100 steps) in a given precision, with / without trace buffers.
usage: python tools/diag_cfg2.py <fp32|bf16|bf16act> <trace 0|1> [rows]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402
from conftest import cfg2_inputs, seed0_state_dict  # noqa: E402


def main():
    prec, trace = sys.argv[1], int(sys.argv[2])
    rows = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    ge.build()
    from sddm_b200 import PREC_BF16, PREC_BF16_ACT, PREC_FP32
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM
    dev = torch.device("cuda:0")
    sd, net = seed0_state_dict()
    net.precision = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16act": PREC_BF16_ACT, "bf16x3": 3}[prec]
    model = SDDM(GaussianDiffusion("linear", 100, 1e-6, 1e-3, device=dev), net, p_transition="condition_in").to(dev).eval()
    cond, noises = cfg2_inputs()
    cond, noises = cond[:rows].to(dev), noises[:, :rows].contiguous().to(dev)
    for rep in range(3):
        res = model.infer(cond, noises=noises, return_trace=bool(trace))
        torch.cuda.synchronize()
        out = res[0] if trace else res
        print("ok", prec, "trace", trace, "rep", rep, float(out.abs().max()), flush=True)


if __name__ == "__main__":
    main()
