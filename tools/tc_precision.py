"""Prints the error of the convolution kernels (fp32 FFMA, tensor-core bf16x3, bf16x6) against an fp64
reference for growing contraction lengths.  Run on the GPU box: python tools/tc_precision.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from emu_backend import EmuBackend  # noqa: E402
from gansynth_b200.kernels import CudaBackend  # noqa: E402

emu = EmuBackend()
g = torch.Generator().manual_seed(0)
print("%-28s %5s %12s %12s" % ("shape", "impl", "max/max", "rms/rms"))
for (n, h, w, ci, co) in [(1, 16, 16, 32, 32), (1, 16, 16, 64, 64), (1, 16, 16, 128, 128), (1, 16, 16, 256, 256), (2, 64, 64, 32, 32)]:
    x = torch.randn(n, h, w, ci, generator=g)
    wt = torch.randn(3, 3, ci, co, generator=g)
    want = emu.conv_c(x.double(), wt.double(), None, 3, 1, 0, 1.0, 0)
    for impl in (2, 3, 5):
        k = CudaBackend()
        k.impl = impl
        got = k.conv_c(x.cuda(), wt.cuda(), None, 3, 1, 0, 1.0, 0).cpu().double()
        d = got - want
        print("%-28s %5d %12.3e %12.3e   mean signed err/rms %.3e" % (
            (n, h, w, ci, co), impl, float(d.abs().max() / want.abs().max()), float(d.pow(2).mean().sqrt() / want.pow(2).mean().sqrt()),
            float(d.mean() / want.pow(2).mean().sqrt())))
# positive-only operands expose a truncation bias in the accumulator
x = torch.rand(1, 16, 16, 256, generator=g)
wt = torch.rand(3, 3, 256, 64, generator=g)
want = emu.conv_c(x.double(), wt.double(), None, 3, 1, 0, 1.0, 0)
for impl in (2, 3, 5):
    k = CudaBackend()
    k.impl = impl
    d = k.conv_c(x.cuda(), wt.cuda(), None, 3, 1, 0, 1.0, 0).cpu().double() - want
    print("positive operands K=2304 impl %d: mean rel err %.3e  max rel err %.3e" % (impl, float((d / want).mean()), float((d / want).abs().max())))
