"""Runs the two spectral kernels at batch 256 a few times (for ncu captures / timing)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gansynth_b200.spectral_ops as sp  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = (0.1 * torch.randn(b, 64000, generator=torch.Generator().manual_seed(0))).cuda()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(6):
    if it == 3:
        ev[0].record()
    lm, inst = sp.convert_to_spectrogram(w, **bench.SPECTRAL)
ev[1].record()
for it in range(3):
    back = sp.convert_to_waveform(lm, inst, **bench.SPECTRAL)
ev[2].record()
torch.cuda.synchronize()
print("batch %d: fwd %.1f us, inv %.1f us" % (b, ev[0].elapsed_time(ev[1]) * 1e3 / 3, ev[1].elapsed_time(ev[2]) * 1e3 / 3))
