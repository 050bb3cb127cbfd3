#!/usr/bin/env python
"""tcgen05.mma issue-rate microbenchmark (M=128, K=16 bf16, SS operands): cycles per MMA vs N."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from sddm_b200 import _lib
torch.zeros(1, device="cuda")
lib = _lib.lib()
GEO = {0: "canonical (SBO 128 B, aligned core matrices)", 1: "conv halo geometry (SBO 160 B, 9 tap offsets)", 2: "x-shifted dense copies (SBO 128 B, aligned)"}
for ctas, nA in ((1, 9), (148, 10)):
    for geo in (0, 1, 2):
        for N in (16, 32, 64, 96, 128, 160, 256):
            v = C.c_float()
            _lib.check(lib.sddm_debug_umma_rate(N, 1800, nA, geo, C.byref(v)))
            print("CTAs=%3d N=%3d %-52s %.1f cycles / MMA  (math floor %d, A+B bytes %d)" % (ctas, N, GEO[geo], v.value, max(8, 128 * N // 256), 4096 + N * 32))
