#!/usr/bin/env python
"""STFT front-end throughput: cfg-5-shaped batch (10 s clips, n_fft 1024, hop 256), CUDA events, achieved GB/s of algorithmic bytes
(read 4 L, write 4 * 513 * frames per clip for the linear spectrogram; 4 * 80 * frames for mel)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from sddm_b200 import prepare_spectrogram as PS
dev = torch.device("cuda:0")
B, L, hop = 64, 160000, 256
x = (0.1 * torch.randn(B, L, generator=torch.Generator().manual_seed(0))).to(dev)
frames = 1 + L // hop
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
for name, tr, nout in (("spectrogram(hamming)+log/clamp", PS.Spectrogram(n_fft=1024, hop_length=hop, window_fn=torch.hamming_window, log_clamp=True), 513),
                       ("mel(80)+log/clamp", PS.MelSpectrogram(n_fft=1024, hop_length=hop, f_min=20.0, f_max=8000.0, n_mels=80, sample_rate=16000, log_clamp=True), 80)):
    for _ in range(3):
        tr(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        tr(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    by = B * (4.0 * L + 4.0 * nout * frames)
    print("%-32s B=%d x %.1f s: %.3f ms, %.0f GB/s algorithmic (%.2f of %.0f GB/s HBM peak), %.0f x real time" %
          (name, B, L / 16000, ms, by / ms / 1e6, by / ms / 1e6 / peak, peak, B * L / 16000 / (ms / 1e3)))
