#!/usr/bin/env python
"""Tiny end-to-end run for compute-sanitizer: 2-step sampling of one chunk in every precision mode + one STFT call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from sddm_b200 import PREC_BF16, PREC_BF16_ACT, PREC_FP32
from sddm_b200 import prepare_spectrogram as PS
from sddm_b200.model.diffusion import GaussianDiffusion
from sddm_b200.model.model import SDDM
from sddm_b200.model.network import UNetModified2
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = UNetModified2(num_samples=16448, res_blocks=1)
model = SDDM(GaussianDiffusion("linear", 2, 1e-4, 5e-2, device=dev), net, p_transition="condition_in").to(dev).eval()
cond = (0.1 * torch.randn(1, 1, 16448)).clamp(-1, 1).to(dev)
for prec in (PREC_FP32, PREC_BF16, PREC_BF16_ACT):
    net.precision = prec
    out = model.infer(cond, seed=1)
    torch.cuda.synchronize()
    print("prec", prec, float(out.abs().max()))
print("stft", float(PS.Spectrogram(n_fft=1024, hop_length=256, window_fn=torch.hamming_window)(cond[0]).max()))
