#!/usr/bin/env python
"""Profiling driver: N denoiser forwards (sddm_eps) + posterior steps of a B-row batch, nothing else.
Used under ncu (see profiles/README.md); also prints CUDA-event ms per forward when run plain."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16act", "bf16x3"])
    ap.add_argument("--trace", action="store_true", help="per-role wait-cycle trace of every tcgen05 conv launch of the last forward")
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
    cond = (0.1 * torch.randn(args.batch, 1, 16448, generator=torch.Generator().manual_seed(1))).clamp(-1, 1).to(dev)
    x = cond.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(args.iters):
        e0.record()
        eps = plan.eps(cond, x, t=100 - it)
        e1.record()
        torch.cuda.synchronize()
        print("forward %d: %.3f ms (B=%d, %s)" % (it, e0.elapsed_time(e1), args.batch, args.precision))
        x = model.diffusion.p_transition(x, 100 - it, eps)
    torch.cuda.synchronize()
    if args.trace:
        import ctypes as C
        import numpy as np
        from sddm_b200 import _lib
        lib = _lib.lib()
        _lib.check(lib.sddm_debug_tc_trace(1, None))
        plan.eps(cond, x, t=50)
        buf = np.zeros((64, 48), dtype=np.int64)
        _lib.check(lib.sddm_debug_tc_trace(0, C.c_void_p(buf.ctypes.data)))
        print("conv#  role: total_cycles | waits...   (epilogue: tmem_full, res_full, store_read, store_read+bar, tiles; mma: tmem_empty, full_a, full_w; tma: raw_empty; xform: raw_full, empty_a)")
        for i in range(64):
            r = buf[i]
            if r[16] == 0:
                continue
            f = lambda v: "%7.1fk" % (v / 1e3)
            print("%2d epi0 %s | %s %s %s %s tiles=%d tmem_ld %s ld..bar %s || mma %s | %s %s %s issue %s commit %s || tma %s | %s || xf0 %s | %s %s hdr %s rounds %s fence %s" % (
                i, f(r[0]), f(r[1]), f(r[2]), f(r[3]), f(r[4]), r[5], f(r[6]), f(r[7]), f(r[16]), f(r[17]), f(r[18]), f(r[19]), f(r[20]), f(r[21]), f(r[24]), f(r[25]),
                f(r[32]), f(r[33]), f(r[34]), f(r[35]), f(r[36]), f(r[37])))
    print("ok", float(x.abs().max()))


if __name__ == "__main__":
    main()
