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
for nA in (1, 9):
    for N in (16, 32, 64, 96, 128, 160, 256):
        v = C.c_float()
        _lib.check(lib.sddm_debug_umma_rate(N, 2000, nA, C.byref(v)))
        print("N=%3d distinct A tiles=%d: %.1f cycles / MMA  (math floor %d, A+B bytes %d)" % (N, nA, v.value, max(8, 128 * N // 256), 4096 + N * 32))
