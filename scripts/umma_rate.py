#!/usr/bin/env python
"""Cycles per tcgen05.mma (128 x N x 16, fp16, cta_group::1) on one SM, nothing else running on it
(csrc/microbench.cu).  mode 0 = A, B from shared memory; 1 = A from tensor memory; 2 = SS + concurrent smem writers;
3 = A copied smem -> TMEM by tcgen05.cp before each TS-mode MMA (+4: issuing thread chosen by elect.sync)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402  (creates the CUDA context)

from audioldm2_b200 import _lib  # noqa: E402

torch.zeros(1, device="cuda")
L = _lib.lib()
n_sm = 148
reps = 2048
print(f"{'mode':>4s} {'N':>4s} {'cycles/mma (median over SMs)':>30s} {'floor N/2':>10s}")
for mode in (0, 4, 5, 6, 7):   # bit 2: issuing thread chosen by elect.sync (0: lane == 0, the pre-fix code); 7 = tcgen05.cp + TS MMA
    for N in (32, 64, 128, 256):
        buf = (C.c_longlong * n_sm)()
        for _ in range(2):
            _lib.check(L.aldm_debug_umma_rate(N, mode, reps, buf, n_sm), "umma_rate")
        v = sorted(buf)
        print(f"{mode:4d} {N:4d} {v[len(v) // 2] / reps:30.1f} {N / 2:10.1f}", flush=True)
