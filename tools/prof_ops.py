#!/usr/bin/env python
"""Per-op CUDA-event timing of the UNetModified2 op program (one line per launch of a forward): label, us, algorithmic GB/s, TFLOP/s.
usage: python tools/prof_ops.py [--batch 64] [--iters 5] [--precision bf16act]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--precision", default="bf16act", choices=["bf16", "fp32", "bf16act", "bf16x3"])
    ap.add_argument("--trace", action="store_true", help="per-role wait-cycle trace of the conv_row launches of one forward")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    ge.build()
    from sddm_b200 import PREC_BF16, PREC_BF16_ACT, PREC_FP32
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM
    from sddm_b200.model.network import UNetModified2
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = UNetModified2(num_samples=16448, res_blocks=1)
    net.precision = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16act": PREC_BF16_ACT, "bf16x3": 3}[args.precision]
    model = SDDM(GaussianDiffusion("linear", 100, 1e-6, 1e-3, device=dev), net, p_transition="condition_in").to(dev).eval()
    plan = net.get_plan(model.diffusion)
    B = args.batch
    cond = (0.1 * torch.randn(B, 1, 16448, generator=torch.Generator().manual_seed(1))).clamp(-1, 1).to(dev)
    x = cond.clone()
    for it in range(3):
        plan.eps(cond, x, t=100)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(args.iters):
        plan.eps(cond, x, t=100 - it)
    e1.record()
    torch.cuda.synchronize()
    print("forward (unprofiled): %.3f ms  (B=%d, %s, %d launches)" % (e0.elapsed_time(e1) / args.iters, B, args.precision, plan.launches_per_eps()))
    plan.profile(True)
    for it in range(args.iters):
        plan.eps(cond, x, t=100 - it)
    torch.cuda.synchronize()
    rep = plan.profile_report()
    plan.profile(False)
    tot = 0.0
    for o in rep:
        if not o["launches"]:
            continue
        us = 1e3 * o["ms"] / o["launches"]
        tot += us
        print("%-22s %8.1f us  %7.0f GB/s  %7.1f TFLOP/s  %s" % (o["label"], us, o["bytes_per_row"] * B / us / 1e3, o["flops_per_row"] * B / us / 1e6,
                                                                "tc" if o["tensor_cores"] else ""))
    print("sum of per-op times: %.3f ms" % (tot / 1e3))
    if args.trace:
        import ctypes as C
        import numpy as np
        from sddm_b200 import _lib
        lib = _lib.lib()
        _lib.check(lib.sddm_debug_row_trace(1, None))
        plan.eps(cond, x, t=50)
        buf = np.zeros((64, 32), dtype=np.int64)
        _lib.check(lib.sddm_debug_row_trace(0, C.c_void_p(buf.ctypes.data)))
        k = lambda v: "%6.1fk" % (v / 1e3)
        print("row kernel trace (CTA 0, kcycles): launch | epi0: total acc_full tmem..sts barrier bulk_wait tma_issue issue+stats rows | mma: total w acc_empty full_a rows | tma: total raw_empty | xf0: total raw_full empty_a work")
        for i in range(64):
            r = buf[i]
            if r[16] == 0:
                continue
            print("%2d | %s %s %s %s %s %s %s %3d | %s %s %s %s %3d | %s %s | %s %s %s %s" % (
                i, k(r[0]), k(r[1]), k(r[2]), k(r[3]), k(r[4]), k(r[5]), k(r[6]), r[7], k(r[16]), k(r[17]), k(r[18]), k(r[19]), r[20],
                k(r[21]), k(r[22]), k(r[24]), k(r[25]), k(r[26]), k(r[27])))


if __name__ == "__main__":
    main()
