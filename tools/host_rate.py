#!/usr/bin/env python
"""Host enqueue rate vs GPU execution rate of the denoiser op program: if the host needs about as long to enqueue a forward (44 launches,
tensor-map encodes included) as the GPU needs to run it, programmatic dependent launch finds the queue empty and every kernel pays its
launch latency.  usage: python tools/host_rate.py [--batch 64]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    ge.build()
    from sddm_b200 import PREC_BF16_ACT
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM
    from sddm_b200.model.network import UNetModified2
    dev = torch.device("cuda:0")
    net = UNetModified2(num_samples=16448, res_blocks=1)
    net.precision = PREC_BF16_ACT
    model = SDDM(GaussianDiffusion("linear", 100, 1e-6, 1e-3, device=dev), net, p_transition="condition_in").to(dev).eval()
    plan = net.get_plan(model.diffusion)
    B = args.batch
    cond = (0.1 * torch.randn(B, 1, 16448, generator=torch.Generator().manual_seed(1))).clamp(-1, 1).to(dev)
    x = cond.clone()
    for _ in range(3):
        plan.eps(cond, x, t=50)
    torch.cuda.synchronize()
    print("cpu affinity:", sorted(os.sched_getaffinity(0))[:8], "... of", os.cpu_count())
    for rep in range(4):
        n = 8
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        t0 = time.perf_counter()
        for _ in range(n):
            plan.eps(cond, x, t=50)
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("rep %d: host enqueue %.3f ms / forward, gpu %.3f ms / forward, wall %.3f ms / forward" % (rep, (t1 - t0) / n * 1e3, e0.elapsed_time(e1) / n, (t2 - t0) / n * 1e3))
    # a full sampling run (what bench.py times)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = model.infer(cond, seed=1)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("infer rep %d: host returned after %.1f ms, done after %.1f ms" % (rep, (t1 - t0) * 1e3, (t2 - t0) * 1e3))


if __name__ == "__main__":
    main()
