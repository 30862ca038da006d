"""Cycles per tcgen05.mma (M=128, K=16, bf16) for K-major vs MN-major SWIZZLE_NONE operands and several N.
Run on the GPU box: python tools/mma_timing.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gansynth_b200 import _lib  # noqa: E402

st = torch.cuda.current_stream().cuda_stream
cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
print("%-9s %4s %4s %8s %10s %12s" % ("layout", "N", "K", "gstride", "cycles/MMA", "MAC/cycle"))
for mode, name in ((0, "K-major"), (1, "MN-major")):
    for n in (32, 64, 128, 256):
        for gstride in (8, 10):
            k = 64
            rows_a = (15 * gstride + 8 + 4) if mode == 0 else ((k // 8 - 1) * gstride + 8 + 4)
            rows_b = n if mode == 0 else rows_a
            a = torch.randn(rows_a, k if mode == 0 else 128, device="cuda")
            b = torch.randn(rows_b, k if mode == 0 else n, device="cuda")
            d = torch.zeros(128, n, device="cuda")
            reps = 2000
            _lib.probe_call("gs_tc_probe_time", a.data_ptr(), b.data_ptr(), d.data_ptr(), k, n, rows_a, rows_b, 0, gstride, mode,
                      reps, cyc.data_ptr(), st)
            torch.cuda.synchronize()
            per = float(cyc.item()) / (reps * (k // 16))
            print("%-9s %4d %4d %8d %10.1f %12.0f" % (name, n, k, gstride, per, 128 * n * 16 / per))
