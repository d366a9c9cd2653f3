#!/usr/bin/env python
"""Pacing of small tcgen05.mma on this GPU: cycles per MMA (issue / completion) for N = 16..256, dependent chains vs
several accumulators, one vs two issuing warps, one vs all SMs."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vanerf_b200.renderer import Renderer  # noqa: E402

r = Renderer("cuda:0")
out = (C.c_longlong * 4)()
REPS = 512
print("   N n_acc warps ctas | issue cyc/mma  total cyc/mma  (floor N/2)")
for n_ctas in (1, 148):
    for mode in (0, 1):
        for n, n_acc in [(16, 1), (16, 4), (32, 1), (64, 1), (64, 2), (64, 4), (96, 1), (128, 1), (128, 2), (256, 1)]:
            st = r.lib.dll.vanerf_tc_mma_probe(r.ctx, n, REPS, n_acc, mode, n_ctas, out)
            if st:
                print("probe failed", st, r.lib.last_error(r.ctx) if hasattr(r.lib, "last_error") else "")
                continue
            print(f"{n:4d} {n_acc:5d} {1 + mode:5d} {n_ctas:4d} | {out[0] / REPS:10.1f} {out[1] / REPS:14.1f}   ({n / 2:.0f})"
                  + (f"   warp1: {out[2] / REPS:.1f} {out[3] / REPS:.1f}" if mode else ""), flush=True)
