#!/usr/bin/env python
"""Which of the two per-kernel latency modes does this process run in, and does it correlate with anything
observable?  Prints the per-forward time at B = 64, the time of a full 100-step run, the graph-replay latency of a 2-chunk clip,
device addresses and clocks.  Variants (environment):
  PROBE_PRE=pin|dev|big   allocations made before the plan exists (pinned host / device / PROBE_BIG_MB MB of device memory)
  PROBE_ARENA_FIRST=1     bench.py's order: warm-up runs, the host-buffer call at 64 rows, then the timed loops
  PROBE_STREAM=1          timed loops on a side stream instead of the legacy default stream
  PROBE_INFER=1, PROBE_PROFILE=1, PROBE_ARENA64=1   extra phases before the clip (full runs, the per-op profiling pass, a 64-row arena)
  PROBE_REALLOC=1         re-allocate the plan's arena ten times and time the clip after each
  PROBE_CLIPS=n           n more clips (drift?), a heavy burst, the clip again
Result of round 2 (40+ processes, with and without ASLR): this script is always in the slow mode; bench.py is not - see DESIGN.md."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import __graft_entry__ as ge
    ge.build()
    from sddm_b200 import PREC_BF16_ACT
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM
    from sddm_b200.model.network import UNetModified2
    dev = torch.device("cuda:0")
    L = 16448
    pre = os.environ.get("PROBE_PRE", "")        # allocations made BEFORE the plan exists (bench.py makes both)
    keep = []
    if "pin" in pre:
        keep.append(torch.zeros(64, 1, L).pin_memory())
        keep.append(torch.zeros(64, 1, L).pin_memory())
    if "dev" in pre:
        keep.append(torch.zeros(64, 1, L, device=dev))
    if "big" in pre:
        keep.append(torch.zeros(int(os.environ.get("PROBE_BIG_MB", "64")) * 262144, device=dev))
    net = UNetModified2(num_samples=16448, res_blocks=1)
    net.precision = PREC_BF16_ACT
    model = SDDM(GaussianDiffusion("linear", 100, 1e-6, 1e-3, device=dev), net, p_transition="condition_in").to(dev).eval()
    plan = net.get_plan(model.diffusion)
    cond = (0.1 * torch.randn(64, 1, L, generator=torch.Generator().manual_seed(1))).clamp(-1, 1).to(dev)
    x = cond.clone()
    if os.environ.get("PROBE_ARENA_FIRST") == "1":   # bench.py's order: warm-up forwards, then the host-buffer call at 64 rows, then the timed loops
        for i in range(3):
            model.infer(cond, seed=i, row0=0)
        if len(keep) >= 2 and not keep[0].is_cuda:   # the pinned buffers made before the plan existed (bench.py's cond_host / out_host)
            keep[0].copy_(cond.cpu())
            plan.enhance_host(keep[0], "condition_in", seed=0, row0=0, max_rows=64, out=keep[1])
        else:
            c64 = cond.cpu().pin_memory()
            plan.enhance_host(c64, "condition_in", seed=0, row0=0, max_rows=64, out=torch.empty_like(c64).pin_memory())
    import contextlib
    side = torch.cuda.Stream() if os.environ.get("PROBE_STREAM") == "1" else None      # a non-default stream instead of the legacy stream 0
    with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
        for _ in range(3):
            plan.eps(cond, x, t=50)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            plan.eps(cond, x, t=50)
        e1.record()
        torch.cuda.synchronize()
        fwd = e0.elapsed_time(e1) / 20
        t0 = time.perf_counter()
        for i in range(3):
            model.infer(cond, seed=i, row0=0)
        torch.cuda.synchronize()
        print("infer x3 (B = 64, 100 steps): %.1f ms each" % ((time.perf_counter() - t0) * 1e3 / 3))
    if os.environ.get("PROBE_INFER") == "1":     # bench.py's resident step: the whole 100-step loop inside the library
        for i in range(3):
            model.infer(cond, seed=i, row0=0)
        torch.cuda.synchronize()
    if os.environ.get("PROBE_PROFILE") == "1":   # bench.py's per-op pass
        plan.profile(True)
        model.infer(cond, seed=9, row0=0)
        torch.cuda.synchronize()
        plan.profile_report()
        plan.profile(False)
    if os.environ.get("PROBE_ARENA64") == "1":   # size the plan's own arena for 64 rows first (what bench.py's e2e leg does)
        c64 = cond.cpu().pin_memory()
        plan.enhance_host(c64, "condition_in", seed=0, row0=0, max_rows=64, out=torch.empty_like(c64).pin_memory())
    c2 = (0.05 * torch.randn(2, 1, L, generator=torch.Generator().manual_seed(1))).pin_memory()
    o2 = torch.empty_like(c2).pin_memory()
    ts = []
    for i in range(6):
        t0 = time.perf_counter()
        plan.enhance_host(c2, "condition_in", seed=i, row0=0, max_rows=2, out=o2)
        ts.append(1e3 * (time.perf_counter() - t0))
    if os.environ.get("PROBE_REALLOC") == "1":   # does the mode follow the placement of the plan's arena? (a larger batch re-allocates it)
        res = []
        for R in (2, 3, 4, 5, 6, 7, 8, 2, 3, 2):
            cR = (0.05 * torch.randn(R, 1, L, generator=torch.Generator().manual_seed(1))).pin_memory()
            oR = torch.empty_like(cR).pin_memory()
            if R == 2:   # force a re-allocation for the repeated sizes too: grow first, which frees the arena
                plan.enhance_host(torch.zeros(9, 1, L).pin_memory(), "condition_in", seed=0, row0=0, max_rows=9, out=torch.zeros(9, 1, L).pin_memory())
            tt = []
            for i in range(5):
                t0 = time.perf_counter()
                plan.enhance_host(cR, "condition_in", seed=i, row0=0, max_rows=R, out=oR)
                tt.append(1e3 * (time.perf_counter() - t0))
            res.append("%d:%.1f" % (R, min(tt[2:])))
        print("rows:latency after re-allocating the arena:", " ".join(res))
    n_more = int(os.environ.get("PROBE_CLIPS", "0"))
    if n_more:   # does the clip latency drift while the GPU only sees this light load?
        more, clk = [], []
        for i in range(n_more):
            t0 = time.perf_counter()
            plan.enhance_host(c2, "condition_in", seed=i, row0=0, max_rows=2, out=o2)
            more.append(1e3 * (time.perf_counter() - t0))
            if i % 25 == 24:
                clk.append(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
        print("clip latencies every 10th:", " ".join("%.1f" % v for v in more[::10]))
        print("clocks while looping:", clk)
        # a heavy burst, then the clip again
        for _ in range(300):
            plan.eps(cond, x, t=50)
        torch.cuda.synchronize()
        after = []
        for i in range(5):
            t0 = time.perf_counter()
            plan.enhance_host(c2, "condition_in", seed=i, row0=0, max_rows=2, out=o2)
            after.append(1e3 * (time.perf_counter() - t0))
        print("clip right after 300 forwards at B = 64:", " ".join("%.1f" % v for v in after))
    q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,clocks.gr,clocks.video,pstate,power.draw,temperature.gpu,temperature.memory",
                        "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    print("pid %d  forward %.3f ms  clip %.1f ms  ws 0x%x  cond 0x%x  cpu %s  | %s" % (
        os.getpid(), fwd, min(ts[2:]), plan.workspace(64).data_ptr(), cond.data_ptr(), sorted(os.sched_getaffinity(0))[:2], q))


if __name__ == "__main__":
    main()
