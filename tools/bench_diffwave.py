#!/usr/bin/env python
"""cfg 5 measurement: DiffWave on 10 s utterances (wav -> STFT front-end -> [513, 626] -> 200-step sampling, T = 160 256).

Prints one JSON object: conditioner time (one-off per batch), per-eps_hat time, the achieved HBM bandwidth of one eps_hat
evaluation against its algorithmic bytes (tcgen05 path, per audio sample: 30 layers x (x in 128 + cached conditioner 256 +
x out 128 + z 128 = 640 B) + head (30 x 128 B of z) + 8 B of audio / eps = 23 048 B; fp32 path: 30 x 2560 + 264 B) and its
tensor throughput (63.2 kFLOP per sample per layer incl. the skip contraction), and full-sampling utterances/s + real-time factor.
`--steps N` limits the sampling loop to the last N steps of the schedule for a quick run (reported as such).
`--cpu-sample-frames F` also times the oracle (CPU port of the reference) on F frames of one utterance, one step.
"""
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
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--eps-iters", type=int, default=5)
    ap.add_argument("--full", action="store_true", help="run the full 200-step sampling (otherwise extrapolate from eps timing)")
    ap.add_argument("--cpu-sample-frames", type=int, default=0)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    ge.build()
    from sddm_b200 import PREC_BF16, PREC_FP32, prepare_spectrogram as PS
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM_spectrogram
    from sddm_b200.model.network import DiffWave
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = DiffWave(freq_bins=513)
    with torch.no_grad():
        net.output_projection.weight.copy_(0.1 * torch.randn(net.output_projection.weight.shape))
    net.precision = {"fp32": PREC_FP32, "bf16": PREC_BF16}[args.precision]
    d = GaussianDiffusion("linear", 200, 1e-4, 0.02, device=dev)
    model = SDDM_spectrogram(d, net, hop_samples=256, noise_condition="time_step").to(dev).eval()
    B, L = args.batch, int(args.seconds * 16000)
    wav = (0.1 * torch.randn(B, L, generator=torch.Generator().manual_seed(1))).to(dev)
    spec = PS.Spectrogram(n_fft=1024, hop_length=256, window_fn=torch.hamming_window, log_clamp=True)(wav).contiguous()
    frames = spec.shape[-1]
    T = 256 * frames
    plan = net.get_plan(d, "time_step")
    ev = lambda: torch.cuda.Event(enable_timing=True)
    out = dict(workload="DiffWave cfg5: %d x %.0f s utterances, spec [513,%d], T=%d, 200 steps, %s" % (B, args.seconds, frames, T, args.precision))
    # conditioner (one-off per batch)
    plan.condition(spec)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record(); plan.condition(spec); e1.record(); torch.cuda.synchronize()
    out["condition_ms"] = e0.elapsed_time(e1)
    # eps_hat
    audio = torch.randn(B, 1, T, device=dev)
    for _ in range(2):
        plan.eps(spec, audio, t=100)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.eps_iters):
        plan.eps(spec, audio, t=100 - i)
    e1.record(); torch.cuda.synchronize()
    eps_ms = e0.elapsed_time(e1) / args.eps_iters
    out["eps_ms"] = eps_ms
    layers = 30
    alg_bytes = B * T * ((640.0 * layers + 128.0 * layers + 8.0) if args.precision == "bf16" else (2560.0 * layers + 264.0))
    alg_flops = B * T * (2.0 * (192 * 128 + 64 * 128) * layers + 2.0 * 64 * 64 + 128)
    out["eps_alg_GBps"] = alg_bytes / eps_ms / 1e6
    out["eps_TFLOPs"] = alg_flops / eps_ms / 1e9
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    out["hbm_peak_GBps"] = peaks["hbm_gbs"]
    out["hbm_frac"] = out["eps_alg_GBps"] / peaks["hbm_gbs"]
    est = (out["condition_ms"] + 200 * eps_ms) / 1e3
    out["sampling_s_estimated"] = est
    out["utt_per_s_estimated"] = B / est
    out["rtf_estimated"] = est / (B * args.seconds)
    if args.full:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record(); x0 = model.infer(spec, seed=1); e1.record(); torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        out["sampling_s"] = e0.elapsed_time(e1) / 1e3
        out["sampling_wall_s"] = wall
        out["utt_per_s"] = B / out["sampling_s"]
        out["rtf"] = out["sampling_s"] / (B * args.seconds)
        out["x0_absmax"] = float(x0.abs().max())
    if args.cpu_sample_frames:
        from oracle import diffwave_oracle as DO
        F = args.cpu_sample_frames
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        sp, au = spec[:1, :, :F].cpu(), audio[:1, :, :256 * F].cpu()
        torch.set_num_threads(os.cpu_count())
        DO.diffwave_forward(sd, sp, au, torch.full((1, 1, 1), 100.0))
        t0 = time.perf_counter()
        DO.diffwave_forward(sd, sp, au, torch.full((1, 1, 1), 100.0))
        dt = time.perf_counter() - t0
        out["cpu_port"] = dict(seconds_per_eps=dt, frames=F, cores=os.cpu_count(), scaled_s_per_utt=dt * frames / F * 200,
                               sample="1 eps_hat on %d of %d frames of one utterance, scaled linearly" % (F, frames))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
