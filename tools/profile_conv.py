"""Runs single convolution layers of the full-size step a few times (for ncu captures).
usage: python tools/profile_conv.py [form] [n h w ci co stride]   form in c|t|w"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gansynth_b200.kernels import CudaBackend  # noqa: E402

form = sys.argv[1] if len(sys.argv) > 1 else "c"
n, h, w, ci, co, st = (int(v) for v in sys.argv[2:8]) if len(sys.argv) > 7 else (8, 128, 1024, 32, 32, 1)
k = CudaBackend()
g = torch.Generator().manual_seed(0)
x = torch.randn(n, h, w, ci, generator=g).cuda()
dy = torch.randn(n, h // st, w // st, co, generator=g).cuda()
wt = torch.randn(3, 3, ci, co, generator=g).cuda()
b = torch.randn(co, generator=g).cuda()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(6):
    if it == 3:
        ev[0].record()
    if form == "c":
        k.conv_c(x, wt, b, 3, st, 0, 0.05, 1)
    elif form == "t":
        k.conv_t(dy, wt, None, 3, st, 0, 0.05, 0)
    else:
        k.conv_w(x, dy, 3, st, 0, 0.05)
ev[1].record()
torch.cuda.synchronize()
print("%s %s: %.1f us per call" % (form, (n, h, w, ci, co, st), ev[0].elapsed_time(ev[1]) * 1e3 / 3))
