#!/usr/bin/env python
"""Headline benchmark: utterances/s (and real-time factor) of FULL reverse-diffusion enhancement with the
UNetModified2 denoiser (config_unet.json: 100 steps, 16448-sample chunks), BASELINE.json configs[1]:
a batch of 64 VoiceBank-DEMAND-shaped utterance-chunks per GPU, synthetic audio, config-shaped random-init weights.

    python bench.py [--gpus N --steps K --warmup W]                 # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference [...]                           # the reference algorithm on the host CPU cores
    python bench.py --workload cfg3 [...]                             # BASELINE.json configs[2]: 824 test-set-shaped utterances, STRONG scaling:
                                                                      # chunk -> shard rows over the ranks -> enhance -> NCCL gather -> regroup
    python bench.py --workload cfg4 [...]                             # BASELINE.json configs[3]: WaveGrad, 32 x 2 s utterances per GPU, 1000 steps
    python bench.py --workload cfg5 [...]                             # BASELINE.json configs[4]: DiffWave, 8 x 10 s utterances per GPU,
                                                                      # 200 steps (same JSON contract; not the headline)

One "step" = one full enhancement (x_T init + 100 x [eps_hat, posterior update]) of the rank's 64-row batch.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L, T_STEPS, SR = 16448, 100, 16000
UNET = dict(num_samples=L, in_channel=2, out_channel=1, inner_channel=32, norm_groups=32, channel_mults=(1, 2, 3, 4, 5),
            res_blocks=1, dropout=0, segment_len=128, segment_stride=64)
WORKLOAD = "cfg2: UNetModified2 full 100-step reverse schedule, 64 utterance-chunks x 16448 samples (1.028 s @16 kHz) per GPU"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]), bf16_sustained=float(p["bf16_tflops_sustained"]),
                    source="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        busy = [v for v in sm if v > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def init_distributed(local):
    """init_process_group + first barrier with file descriptor 1 pointing at stderr: NCCL prints its version banner (NCCL_DEBUG >=
    VERSION, possibly from /etc/nccl.conf) on STDOUT when the first communicator is created, next to the one JSON line this script owes."""
    import torch
    import torch.distributed as dist
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def synth_batch(rows, seed):
    import torch
    return (0.1 * torch.randn(rows, 1, L, generator=torch.Generator().manual_seed(seed))).clamp(-1, 1)


# ------------------------------------------------------------------------------------------------------
# reference arms.  The reference is pure Python (no setup.py, nothing to pip-install): __graft_entry__.build() copies its import
# closure (model/ base/ utils/ logger/) to baseline/_ref/ (git-ignored, travels to the GPU box); when that copy is present the
# arms below run the reference's OWN modules (kind "reference"), else the oracle port (same ATen kernels, kind "port").
# ------------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_sddm(n_timestep, device):
    """(kind, model): SDDM(GaussianDiffusion(linear 1e-6..1e-3, n_timestep), UNetModified2(config_unet.json), 'condition_in') built
    from the reference's own classes (reference infer.py:36-42), default init under torch.manual_seed(0)."""
    import torch
    if os.path.isdir(os.path.join(REF_DIR, "model")):
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import model.diffusion as rd
        import model.model as rm
        import model.network as rn
        torch.manual_seed(0)
        net = rn.UNetModified2(**UNET)
        diff = rd.GaussianDiffusion(schedule="linear", n_timestep=n_timestep, linear_start=1e-6, linear_end=1e-3, device=device)
        return "reference", rm.SDDM(diff, net, p_transition="condition_in").to(device).eval()
    return "port", None


class _PortSDDM:
    """oracle port of SDDM.infer for a k-step schedule (used only when baseline/_ref is absent)."""

    def __init__(self, n_timestep):
        import torch
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import sddm_oracle as O
        from sddm_b200.model.network import UNetModified2
        torch.manual_seed(0)
        net = UNetModified2(**UNET)
        self.O, self.T = O, n_timestep
        self.sd = {"noise_estimate_model." + k: v.detach() for k, v in net.state_dict().items()}
        self.sch = O.make_schedule("linear", n_timestep, 1e-6, 1e-3)

    def infer(self, cond):
        import torch
        O = self.O
        x = O.get_x_T(self.sch, self.T, cond, torch.randn_like(cond))
        for t in range(self.T, 0, -1):
            eps = O.unet_forward(self.sd, dict(UNET), cond, x, self.sch["sqrt_alpha_bar"][t] * torch.ones(cond.shape[0], 1, 1))
            x = O.p_transition(self.sch, x, t, eps, torch.randn_like(cond), "condition_in")
        return x


def cpu_reference(rows, k, n_runs, warmup=0, threads=None):
    """Times `n_runs` CPU runs of the reference's SDDM.infer on the real `rows`-chunk batch with a k-step schedule (every reverse
    step costs the same: k of the 100 steps is a bounded sample of the workload).  Nothing is extrapolated over rows."""
    import torch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    kind, model = reference_sddm(k, "cpu")
    if model is None:
        model = _PortSDDM(k)
    cond = synth_batch(rows, 1000)
    times = []
    with torch.no_grad():
        for i in range(warmup + n_runs):
            t0 = time.perf_counter()
            model.infer(cond)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return dict(kind=kind, times=times, rows=rows, k=k, threads=threads)


def cpu_probe_step_seconds(rows):
    """seconds of one reverse step on `rows` chunks (one 1-step run after a warm-up run)."""
    r = cpu_reference(rows, 1, 1, warmup=1)
    return r["times"][0], r["kind"]


def run_reference(args):
    """bench.py --impl reference: the reference's own CPU implementation on all host cores, on the SAME 64-chunk batch; one bench step
    = SDDM.infer with k reverse steps, k sized so that (steps + warmup) runs end within ~150 s; value = rows / (t * 100 / k)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = args.batch
    per, kind = cpu_probe_step_seconds(rows)
    n = max(1, args.steps) + max(0, args.warmup)
    k = max(1, min(T_STEPS, int(args.ref_budget / (n * max(per, 1e-3)))))
    r = cpu_reference(rows, k, max(1, args.steps), warmup=max(0, args.warmup))
    ms = 1e3 * statistics.mean(r["times"])
    v = rows / (ms / 1e3 * T_STEPS / k)
    sample = "%d chunks (the real batch) x %d of 100 reverse steps per bench step through SDDM.infer (k-step schedule); utt/s = rows / (t x 100 / %d)" % (rows, k, k)
    line = {"impl": "reference", "metric": "utterances_per_sec", "value": v, "unit": "utt/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows": rows, "reverse_steps_run": k, "reverse_steps_full": T_STEPS,
                       "note": ("the reference's own modules (baseline/_ref copy of /root/reference)" if kind == "reference" else
                                "oracle port of the reference (torch CPU ATen kernels)") + " on the host cores; ms_per_step is the measured time of one bounded bench step"},
            "rtf": 1.0 / (v * L / SR),
            "cpu_baseline": {"value": v, "unit": "utt/s", "cores": r["threads"], "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def gpu_eager_reference(rows, runs=3, warmup=1, device="cuda:0"):
    """The reference's own PyTorch path on THIS GPU ("the kernel to beat"): reference modules .to(cuda), cudnn.benchmark = True
    (reference infer.py:17), default TF32 convolutions, full 100-step SDDM.infer on the same batch; CUDA events, 1 warm-up + 3 runs."""
    import torch
    kind, model = reference_sddm(T_STEPS, device)
    if model is None:
        return None
    torch.backends.cudnn.benchmark = True
    cond = synth_batch(rows, 1000).to(device)
    times = []
    with torch.no_grad():
        for i in range(warmup + runs):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            model.infer(cond)
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                times.append(e0.elapsed_time(e1))
    best, med = min(times), statistics.median(times)
    del model
    torch.cuda.empty_cache()
    return {"value": rows / (med / 1e3), "unit": "utt/s", "best": rows / (best / 1e3), "ms_median": med, "ms_best": best, "runs": runs, "warmup": warmup,
            "rows": rows, "kind": "reference modules, PyTorch eager + cuDNN (cudnn.benchmark=True, conv TF32 = %s), fp32 tensors"
                                  % torch.backends.cudnn.allow_tf32, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}


def run_reference_gpu(args):
    """bench.py --impl reference-gpu: the same JSON contract, measured on the reference's CUDA-eager path (one GPU)."""
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    g = gpu_eager_reference(args.batch, runs=max(1, args.steps), warmup=max(1, args.warmup))
    if g is None:
        print(json.dumps({"impl": "reference-gpu", "unavailable": "baseline/_ref (copy of the reference modules) is absent"}))
        return
    line = {"impl": "reference-gpu", "metric": "utterances_per_sec", "value": g["value"], "unit": "utt/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": g["ms_median"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "note": g["kind"]}, "rtf": 1.0 / (g["value"] * L / SR), "gpu_eager_baseline": g,
            "e2e": {"value": g["value"], "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# cfg 3: the 824-utterance VoiceBank-DEMAND-test-shaped set, sharded over the ranks (strong scaling)
# ------------------------------------------------------------------------------------------------------
def cfg3_lengths():
    """824 utterance lengths, log-normal with a 2.3 s median clipped to [1 s, 10 s] at 16 kHz, seeded (the real
    test set is not available offline; this reproduces its shape: 824 files, ~2.5 s average)."""
    import numpy as np
    rng = np.random.default_rng(824)
    sec = np.clip(rng.lognormal(mean=np.log(2.3), sigma=0.45, size=824), 1.0, 10.0)
    return (sec * SR).astype(np.int64)


def run_cfg3(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        init_distributed(local)
    from sddm_b200 import PREC_BF16, PREC_BF16_ACT, PREC_FP32, _lib
    from sddm_b200.infer import enhance_utterances
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM
    from sddm_b200.model.network import UNetModified2
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    net = UNetModified2(**UNET)
    net.precision = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16act": PREC_BF16_ACT, "bf16x3": 3}[args.precision]
    model = SDDM(GaussianDiffusion("linear", T_STEPS, 1e-6, 1e-3, device=dev), net, p_transition="condition_in").to(dev).eval()
    lengths = cfg3_lengths()
    g = torch.Generator().manual_seed(824)
    waves = [(0.1 * torch.randn(int(n), generator=g)).clamp(-1, 1).pin_memory() for n in lengths]     # host buffers: H2D is inside the step
    n_chunks = int(sum((int(n) + L - 1) // L for n in lengths))
    audio_s = float(lengths.sum()) / SR

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    outs = [None]

    def step(s):
        outs[0] = enhance_utterances(model, waves, batch_chunks=args.batch, seed=s, rank=rank, world=world)

    def timed(steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            step(s)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for s in range(max(3, args.warmup)):
        step(s)
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _lib.launch_count()
    ms = timed(args.steps)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ok = len(outs[0]) == len(lengths) and all(o.shape[-1] == int(n) for o, n in zip(outs[0], lengths))
    phases = {}
    sync_all()
    t_ph = time.perf_counter()
    enhance_utterances(model, waves, batch_chunks=args.batch, seed=0, rank=rank, world=world, timings=phases)   # one extra, untimed step with a phase timeline
    sync_all()
    phases["step_wall_ms"] = 1e3 * (time.perf_counter() - t_ph)
    ph_all = [None] * world
    if world > 1:
        dist.all_gather_object(ph_all, phases)
    else:
        ph_all = [phases]
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    value = len(lengths) * args.steps / (ms / 1e3)
    line = {"metric": "utterances_per_sec", "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "fp32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": "cfg3: UNetModified2 full 100-step enhancement of 824 test-set-shaped utterances (%.0f s of audio, %d chunks of 16448 "
                                   "samples), rows sharded contiguously over the ranks, NCCL all-gather of the enhanced rows, regrouped per utterance"
                                   % (audio_s, n_chunks),
                       "sub_batch_chunks": args.batch, "reverse_steps": T_STEPS, "precision": args.precision, "noise": "in-kernel Philox4x32-10, keyed by the global row",
                       "lengths": "log-normal, median 2.3 s, clipped to [1, 10] s, numpy default_rng(824)",
                       "l2": "per-step working set of a 64-chunk sub-batch >> 126 MB L2; no flush needed"},
            "rtf": (ms / args.steps / 1e3) / audio_s, "chunks_per_sec": n_chunks * args.steps / (ms / 1e3),
            "e2e": {"value": value, "unit": "utt/s", "h2d_bytes_per_step": n_chunks * L * 4 // world, "d2h_bytes_per_step": 0,
                    "note": "the timed step is already end to end from pinned host waveforms (chunking, H2D, enhancement, gather, regroup); outputs stay on the device"},
            "gpu_launches": int(launches), "clocks": clocks, "outputs_ok": bool(ok),
            "timeline": {"note": "one extra step with device syncs between phases, per rank (ms): host chunking of the rank's own rows, H2D + enhancement, "
                                 "NCCL all-gather of the rows, per-utterance regroup", "per_rank": ph_all}}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# cfg 4: WaveGrad (config_wavegrad.json), spec [128, 107] -> 32 100 samples, 1000 reverse steps
# ------------------------------------------------------------------------------------------------------
WG_FRAMES, WG_STEPS, WG_SECONDS = 107, 1000, 107 * 300 / 16000.0
WG_WORKLOAD = "cfg4: WaveGrad (config_wavegrad.json) full 1000-step sampling, %d utterances x 2.0 s (mel spec [128,107], 32100 samples) per GPU"


def wg_cpu_rate(threads=None):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import wavegrad_oracle as WO
    from sddm_b200.model.network import WaveGrad
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = WaveGrad()
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    spec, audio = torch.rand(1, 128, WG_FRAMES, generator=g), torch.randn(1, 300 * WG_FRAMES, generator=g)
    with torch.no_grad():
        WO.wavegrad_forward(sd, spec, audio, torch.tensor([0.5]))
        t0 = time.perf_counter()
        n = 0
        while n < 3 or time.perf_counter() - t0 < 8.0:
            WO.wavegrad_forward(sd, spec, audio, torch.tensor([0.5]))
            n += 1
        per = (time.perf_counter() - t0) / n
    return dict(value=1.0 / (per * WG_STEPS), threads=threads, sample="%d eps_hat evaluations of one utterance, scaled to 1000 steps" % n)


def run_reference_wavegrad(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    vals = [wg_cpu_rate() for _ in range(max(1, min(args.steps, 5)))]
    v = statistics.median(r["value"] for r in vals)
    r = vals[-1]
    line = {"impl": "reference", "metric": "utterances_per_sec", "value": v, "unit": "utt/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * args.batch_cfg5 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WG_WORKLOAD % args.batch_cfg5, "note": "reference algorithm (oracle port, torch CPU ATen kernels) on host cores"},
            "rtf": 1.0 / (v * WG_SECONDS),
            "cpu_baseline": {"value": v, "unit": "utt/s", "cores": r["threads"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": v, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def run_wavegrad(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        init_distributed(local)
    from sddm_b200 import PREC_BF16, PREC_FP32, _lib
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM_spectrogram
    from sddm_b200.model.network import WaveGrad
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    net = WaveGrad()
    prec = "fp32" if args.precision == "fp32" else "bf16"
    net.precision = PREC_FP32 if prec == "fp32" else PREC_BF16
    d = GaussianDiffusion("linear", WG_STEPS, 1e-6, 1e-2, device=dev)
    model = SDDM_spectrogram(d, net, hop_samples=300).to(dev).eval()
    B, Ls = args.batch_cfg5, 300 * WG_FRAMES
    spec_host = torch.rand(B, 128, WG_FRAMES, generator=torch.Generator().manual_seed(3000 + rank)).pin_memory()
    spec = spec_host.to(dev)
    out_host = torch.empty(B, 1, Ls).pin_memory()
    plan = net.get_plan(d)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident(s):
        model.infer(spec, seed=s, row0=rank * B)

    def step_e2e(s):
        x = spec_host.to(dev, non_blocking=True)
        y = model.infer(x, seed=s, row0=rank * B)
        out_host.copy_(y, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for s in range(max(3, args.warmup)):
        step_resident(s)
    step_e2e(0)
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    prof = None
    if rank == 0 and world == 1 and prec == "bf16":
        plan.profile(True)
        timed(step_resident, 1)
        prof = plan.profile_report()
        plan.profile(False)
    ms_e2e = timed(step_e2e, args.steps)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    total = B * world * args.steps
    value, e2e = total / (ms / 1e3), total / (ms_e2e / 1e3)
    pk = peaks()
    line = {"metric": "utterances_per_sec", "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": prec, "data": "synthetic",
            "config": {"workload": WG_WORKLOAD % B, "batch_per_gpu": B, "reverse_steps": WG_STEPS, "noise": "in-kernel Philox4x32-10",
                       "weights": "config-shaped random init (torch.manual_seed(0))", "precision": prec, "noise_condition": "sqrt_alpha_bar",
                       "l2": "activations of one eps_hat (~%.0f MB) >> 126 MB L2; no flush needed" % (B * 180.0)},
            "rtf": 1.0 / (value * WG_SECONDS),
            "e2e": {"value": e2e, "unit": "utt/s", "h2d_bytes_per_step": B * 128 * WG_FRAMES * 4, "d2h_bytes_per_step": B * Ls * 4,
                    "rtf": 1.0 / (e2e * WG_SECONDS)},
            "gpu_launches": int(launches), "clocks": clocks, "peaks": pk["source"]}
    if prof and prof[1]:
        tot_ms, n, flops, byts = prof
        dur = tot_ms * 1e-3 / n
        line["kernels"] = {"wg_conv_tc": {"avg_us": dur * 1e6, "launches": n, "tflops": flops / (tot_ms * 1e-3) / 1e12,
                                          "gbs": byts / (tot_ms * 1e-3) / 1e9, "share": tot_ms / (ms / args.steps)}}
        line["roofline"] = {"kernel": "wg_conv_tc_kernel (all %d conv launches of one sampling run)" % n, "bound": "tensor",
                            "achieved": flops / (tot_ms * 1e-3) / 1e12, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                            "frac": flops / (tot_ms * 1e-3) / 1e12 / pk["bf16_sustained"], "traffic": None, "avg_launch_us": dur * 1e6,
                            "share_of_step": tot_ms / (ms / args.steps), "executed_flops_per_launch": flops / n,
                            "algorithmic_bytes_per_launch": byts / n, "arith_intensity_flop_per_byte": flops / max(byts, 1.0),
                            "note": "flops = GEMM work executed (polyphase up-sampling convs include their zero blocks)"}
    if world == 1 and not args.no_cpu_baseline:
        r = wg_cpu_rate()
        line["cpu_baseline"] = {"value": r["value"], "unit": "utt/s", "cores": r["threads"], "kind": "port", "sample": r["sample"]}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# cfg 5: DiffWave (config_diffwave.json) on 10 s utterances, spectrogram from the STFT front-end, 200 reverse steps
# ------------------------------------------------------------------------------------------------------
DW_SECONDS, DW_STEPS, DW_FRAMES = 10.0, 200, 626
DW_WORKLOAD = ("cfg5: DiffWave (config_diffwave.json) full 200-step sampling, %d utterances x 10 s (spec [513,626] from the STFT "
               "front-end, 160256 samples) per GPU")


def dw_cpu_rate(frames=40, threads=None):
    """utterances/s of the CPU port (oracle) from a bounded sample: one eps_hat on `frames` of 626 frames of one utterance,
    scaled linearly in time (the network is fully convolutional) and to 200 steps."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import diffwave_oracle as DO
    from sddm_b200.model.network import DiffWave
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = DiffWave(freq_bins=513)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    spec, audio = torch.rand(1, 513, frames, generator=g) * 0.7, torch.randn(1, 1, 256 * frames, generator=g)
    with torch.no_grad():
        DO.diffwave_forward(sd, spec, audio, torch.full((1, 1, 1), 100.0))
        t0 = time.perf_counter()
        n = 0
        while n < 3 or time.perf_counter() - t0 < 5.0:
            DO.diffwave_forward(sd, spec, audio, torch.full((1, 1, 1), 100.0 - n))
            n += 1
        per = (time.perf_counter() - t0) / n
    full = per * DW_FRAMES / frames * DW_STEPS
    return dict(value=1.0 / full, seconds_full=full, threads=threads,
                sample="%d eps_hat evaluations on %d of %d frames of one utterance, scaled to 626 frames x 200 steps" % (n, frames, DW_FRAMES))


def run_reference_diffwave(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    vals = [dw_cpu_rate() for _ in range(max(1, min(args.steps, 5)))]
    v = statistics.median(r["value"] for r in vals)
    r = vals[-1]
    line = {"impl": "reference", "metric": "utterances_per_sec", "value": v, "unit": "utt/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * args.batch_cfg5 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "config": {"workload": DW_WORKLOAD % args.batch_cfg5, "note": "reference algorithm (oracle port, torch CPU ATen kernels) on host cores"},
            "rtf": 1.0 / (v * DW_SECONDS),
            "cpu_baseline": {"value": v, "unit": "utt/s", "cores": r["threads"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": v, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def run_diffwave(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        init_distributed(local)
    from sddm_b200 import PREC_BF16, PREC_FP32, _lib, prepare_spectrogram as PS
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM_spectrogram
    from sddm_b200.model.network import DiffWave
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    net = DiffWave(freq_bins=513)
    with torch.no_grad():   # the reference zero-initialises this weight (diffwave.py:131): make eps_hat depend on the network
        net.output_projection.weight.copy_(0.1 * torch.randn(net.output_projection.weight.shape))
    prec = "fp32" if args.precision == "fp32" else "bf16"
    net.precision = PREC_FP32 if prec == "fp32" else PREC_BF16
    d = GaussianDiffusion("linear", DW_STEPS, 1e-4, 0.02, device=dev)
    model = SDDM_spectrogram(d, net, hop_samples=256, noise_condition="time_step").to(dev).eval()
    B, Lw = args.batch_cfg5, int(DW_SECONDS * SR)
    wav = (0.1 * torch.randn(B, Lw, generator=torch.Generator().manual_seed(2000 + rank))).to(dev)
    spec = PS.Spectrogram(n_fft=1024, hop_length=256, window_fn=torch.hamming_window, log_clamp=True)(wav).contiguous()
    frames, Ls = spec.shape[-1], 256 * spec.shape[-1]
    spec_host = spec.cpu().pin_memory()
    out_host = torch.empty(B, 1, Ls).pin_memory()
    plan = net.get_plan(d, "time_step")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident(s):
        model.infer(spec, seed=s, row0=rank * B)

    def step_e2e(s):
        x = spec_host.to(dev, non_blocking=True)                        # H2D of the spectrograms from pinned memory
        y = model.infer(x, seed=s, row0=rank * B)
        out_host.copy_(y, non_blocking=True)                            # D2H of the generated waveforms
        torch.cuda.current_stream().synchronize()

    for s in range(max(3, args.warmup)):
        step_resident(s)
    step_e2e(0)
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    prof = None
    if rank == 0 and world == 1 and prec == "bf16":
        plan.profile(True)
        timed(step_resident, 1)
        prof = plan.profile_report()
        plan.profile(False)
    ms_e2e = timed(step_e2e, args.steps)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    total = B * world * args.steps
    value, e2e = total / (ms / 1e3), total / (ms_e2e / 1e3)
    pk = peaks()
    line = {"metric": "utterances_per_sec", "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": prec, "data": "synthetic",
            "config": {"workload": DW_WORKLOAD % B, "batch_per_gpu": B, "reverse_steps": DW_STEPS, "noise": "in-kernel Philox4x32-10",
                       "weights": "config-shaped random init (torch.manual_seed(0)); output_projection.weight 0.1 N(0,1) instead of zeros",
                       "precision": prec, "noise_condition": "time_step",
                       "l2": "per-layer working set (%.0f MB) >> 126 MB L2; no flush needed" % (B * Ls * 640 / 1e6),
                       "conditioner": "upsampler + 30 conditioner projections evaluated once per step of this bench (inside the timed region) and cached"},
            "rtf": 1.0 / (value * DW_SECONDS),
            "e2e": {"value": e2e, "unit": "utt/s", "h2d_bytes_per_step": B * 513 * frames * 4, "d2h_bytes_per_step": B * Ls * 4,
                    "rtf": 1.0 / (e2e * DW_SECONDS)},
            "gpu_launches": int(launches), "clocks": clocks, "peaks": pk["source"]}
    if prof and prof["layer"][1]:
        lay_ms, lay_n = prof["layer"]
        head_ms, head_n = prof["head"]
        dur = lay_ms * 1e-3 / lay_n
        alg = B * Ls * 640.0
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "top_kernel_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("diffwave_batch") == B:
                traffic = tj.get("diffwave_layer_dram_bytes_per_launch")
        except Exception:
            pass
        line["kernels"] = {"dw_layer_tc": {"avg_us": dur * 1e6, "launches": lay_n, "gbs": alg / dur / 1e9, "share": lay_ms / (ms / args.steps)},
                           "dw_final_tc": {"avg_us": 1e3 * head_ms / max(1, head_n), "launches": head_n,
                                           "gbs": B * Ls * (128.0 * 30 + 4) / (head_ms * 1e-3 / max(1, head_n)) / 1e9, "share": head_ms / (ms / args.steps)}}
        line["roofline"] = {"kernel": "dw_layer_tc_kernel", "bound": "hbm", "achieved": alg / dur / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                            "frac": alg / dur / 1e9 / pk["hbm"], "traffic": traffic, "avg_launch_us": dur * 1e6,
                            "share_of_step": lay_ms / (ms / args.steps), "algorithmic_bytes_per_launch": alg,
                            "algorithmic_bytes_per_sample": 640, "arith_intensity_flop_per_byte": 2.0 * (192 * 128 + 64 * 64) / 640.0}
    if world == 1 and not args.no_cpu_baseline:
        r = dw_cpu_rate()
        line["cpu_baseline"] = {"value": r["value"], "unit": "utt/s", "cores": r["threads"], "kind": "port", "sample": r["sample"]}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        init_distributed(local)
    from sddm_b200 import PREC_BF16, PREC_BF16_ACT, PREC_FP32, _lib
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM
    from sddm_b200.model.network import UNetModified2
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    net = UNetModified2(**UNET)
    net.precision = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16act": PREC_BF16_ACT, "bf16x3": 3}[args.precision]
    model = SDDM(GaussianDiffusion("linear", T_STEPS, 1e-6, 1e-3, device=dev), net, p_transition="condition_in").to(dev).eval()
    B = args.batch
    cond_host = synth_batch(B, 1000 + rank).pin_memory()
    out_host = torch.empty_like(cond_host).pin_memory()
    cond = cond_host.to(dev)
    plan = net.get_plan(model.diffusion)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident(s):
        model.infer(cond, seed=s, row0=rank * B)

    def step_e2e(s):
        # the reference-facing C-ABI call with HOST buffers (sddm_enhance_host): H2D from pinned memory, the full loop and the
        # D2H of the enhanced batch all happen inside the library, inside the timed region; it returns after its stream has drained
        plan.enhance_host(cond_host, "condition_in", seed=s, row0=rank * B, max_rows=B, out=out_host)

    for s in range(max(3, args.warmup)):
        step_resident(s)
    step_e2e(0)
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _lib.launch_count()
    ms = timed(step_resident, args.steps)                               # the headline number: no per-launch events
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    prof = None
    if rank == 0 and world == 1:                                        # separate pass: per-kernel CUDA-event timing
        plan.profile(True)
        timed(step_resident, 1)
        prof = plan.profile_report()
        plan.profile(False)
    ms_e2e = timed(step_e2e, args.steps)
    lat = None
    if rank == 0 and world == 1:
        # BASELINE configs[0] (the reference's CPU-runnable case) on the GPU: ONE 2 s clip = 2 chunks, full 100 steps, end to end through
        # the host-buffer C-ABI call; the latency a single-utterance caller sees (this size is launch / latency bound, not bandwidth bound)
        clip = (0.05 * torch.randn(1, 32000, generator=torch.Generator().manual_seed(1)))
        c2 = torch.nn.functional.pad(clip, (0, 2 * L - 32000)).view(2, 1, L).contiguous().pin_memory()
        o2 = torch.empty_like(c2).pin_memory()
        ts = []
        for i in range(7):
            t0 = time.perf_counter()
            plan.enhance_host(c2, "condition_in", seed=i, row0=0, max_rows=2, out=o2)
            ts.append(1e3 * (time.perf_counter() - t0))
        lat = {"ms_median": statistics.median(ts[2:]), "ms_min": min(ts[2:]), "rows": 2, "clip_seconds": 2.0, "runs": 5, "warmup": 2,
               "rtf": statistics.median(ts[2:]) / 1e3 / 2.0, "path": "sddm_enhance_host (H2D + 100 steps + D2H), wall clock"}
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    total_rows = B * world * args.steps
    value = total_rows / (ms / 1e3)
    e2e = total_rows / (ms_e2e / 1e3)
    pk = peaks()
    line = {"metric": "utterances_per_sec", "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "fp32", "bf16x3": "bf16x3"}.get(args.precision, "bf16"), "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "reverse_steps": T_STEPS, "noise": "in-kernel Philox4x32-10",
                       "weights": "config-shaped random init (torch.manual_seed(0))", "precision": args.precision,
                       "l2": "per-step working set (20-40 MB of activations per chunk, x64 chunks) >> 126 MB L2; no flush needed",
                       "activations": {"fp32": "fp32", "bf16": "fp32 in HBM, bf16 tensor-core operands", "bf16act": "bf16 in HBM, bf16 tensor-core operands, fp32 accumulate",
                                       "bf16x3": "fp32 in HBM, split-bf16 (hi, lo) tensor-core operands, 3 products per MMA step, fp32 accumulate"}[args.precision],
                       "utterance": "one 16448-sample chunk (1.028 s); a 2 s clip is 2 chunks"},
            "rtf": 1.0 / (value * L / SR), "chunks_per_sec": value,
            "e2e": {"value": e2e, "unit": "utt/s", "h2d_bytes_per_step": B * L * 4, "d2h_bytes_per_step": B * L * 4,
                    "rtf": 1.0 / (e2e * L / SR)},
            "gpu_launches": int(launches), "clocks": clocks, "peaks": pk["source"]}
    if lat:
        line["cfg1_latency"] = lat
    if prof:
        tot = sum(o["ms"] for o in prof) or 1.0
        fam = {}
        for o in prof:
            if not o["launches"]:
                continue
            k = o["kernel"] if o["kernel"] != "cuda_core" else o["label"].split(":")[0]
            f = fam.setdefault(k, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
            f["ms"] += o["ms"]; f["launches"] += o["launches"]
            f["flops"] += o["flops_per_row"] * B * o["launches"]; f["bytes"] += o["bytes_per_row"] * B * o["launches"]
        # per kernel FUNCTION: share of the step, average launch, algorithmic GB/s and TFLOP/s over all its launches, and both roofline fractions
        kernels = {k: {"share": f["ms"] / tot, "avg_us": 1e3 * f["ms"] / max(1, f["launches"]), "launches": f["launches"],
                       "tflops": f["flops"] / (f["ms"] * 1e-3) / 1e12, "gbs": f["bytes"] / (f["ms"] * 1e-3) / 1e9,
                       "frac_hbm": f["bytes"] / (f["ms"] * 1e-3) / 1e9 / pk["hbm"], "frac_bf16_sustained": f["flops"] / (f["ms"] * 1e-3) / 1e12 / pk["bf16_sustained"]}
                   for k, f in fam.items()}
        line["kernels"] = kernels
        line["ops"] = {o["label"]: {"us": 1e3 * o["ms"] / o["launches"], "kernel": o["kernel"], "gbs": o["bytes_per_row"] * B * o["launches"] / (o["ms"] * 1e-3) / 1e9,
                                    "tflops": o["flops_per_row"] * B * o["launches"] / (o["ms"] * 1e-3) / 1e12} for o in prof if o["launches"]}
        # roofline block: the dominant kernel = the single op with the largest share of the step (one launch per step)
        top = max((o for o in prof if o["launches"]), key=lambda o: o["ms"])
        dur = top["ms"] * 1e-3 / top["launches"]
        ai = top["flops_per_row"] / max(top["bytes_per_row"], 1.0)
        hbm_bound = ai * pk["hbm"] * 1e9 < pk["bf16_sustained"] * 1e12 or not top["tensor_cores"]
        if hbm_bound:
            ach, peak, unit = top["bytes_per_row"] * B / dur / 1e9, pk["hbm"], "GB/s"
        else:
            ach, peak, unit = top["flops_per_row"] * B / dur / 1e12, pk["bf16_sustained"], "TFLOP/s"
        traffic = None
        try:                                                            # dram bytes of the same kernel from the committed ncu capture
            with open(os.path.join(ROOT, "profiles", "top_kernel_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("batch") == B:
                traffic = tj.get("dram_bytes_per_launch", {}).get(top["label"])
        except Exception:
            pass
        line["roofline"] = {"kernel": "%s (%s)" % (top["kernel"], top["label"]), "bound": "hbm" if hbm_bound else "tensor", "achieved": ach, "peak": peak,
                            "unit": unit, "frac": ach / peak, "traffic": traffic, "avg_launch_us": dur * 1e6,
                            "share_of_step": top["ms"] / tot, "arith_intensity_flop_per_byte": ai,
                            "algorithmic_bytes_per_launch": top["bytes_per_row"] * B,
                            "algorithmic_flops_per_launch": top["flops_per_row"] * B,
                            "family": kernels.get(top["kernel"]),
                            "note": "one launch per step of the dominant kernel function; 'family' = the same figures over ALL launches of that function; "
                                    "'kernels' lists every kernel function, 'ops' every launch of a step"}
    if world == 1 and not args.no_gpu_baseline:
        del model, plan
        torch.cuda.empty_cache()
        line["gpu_eager_baseline"] = gpu_eager_reference(B, device="cuda:%d" % local)
    if world == 1 and not args.no_cpu_baseline:
        per, _ = cpu_probe_step_seconds(B)
        k = max(1, min(T_STEPS, int(args.cpu_budget / max(per, 1e-3))))
        r = cpu_reference(B, k, 1)
        v = B / (r["times"][0] * T_STEPS / k)
        line["cpu_baseline"] = {"value": v, "unit": "utt/s", "cores": r["threads"], "kind": r["kind"],
                                "sample": "%d chunks (the real batch) x %d of 100 reverse steps through SDDM.infer with a %d-step schedule "
                                          "(%.1f s of CPU work); utt/s = rows / (t x 100 / %d)" % (B, k, k, r["times"][0], k)}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--precision", default=os.environ.get("SDDM_B200_PRECISION", "bf16act"), choices=["bf16", "fp32", "bf16act", "bf16x3"])
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference's CUDA-eager run (gpu_eager_baseline)")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: seconds of CPU work for all (steps + warmup) bench steps")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="cfg2 = the headline (UNetModified2, weak scaling); cfg3 = 824-utterance set (strong scaling); cfg4 = WaveGrad; cfg5 = DiffWave")
    ap.add_argument("--batch-cfg5", type=int, default=None, help="utterances per GPU for --workload cfg4 (default 32) / cfg5 (default 8)")
    args = ap.parse_args()
    if args.batch_cfg5 is None:
        args.batch_cfg5 = 32 if args.workload == "cfg4" else 8
    if args.workload == "cfg3" and args.impl == "ours":
        run_cfg3(args)
    elif args.workload == "cfg4":
        (run_reference_wavegrad if args.impl == "reference" else run_wavegrad)(args)
    elif args.workload == "cfg5":
        (run_reference_diffwave if args.impl == "reference" else run_diffwave)(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
