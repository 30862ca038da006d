"""BASELINE config 5: pitch-conditional inference z + pitch -> 128x1024 mel+IF -> 64000-sample audio, batch sweep.
Batches above CHUNK clips run as consecutive generate_batch calls of CHUNK clips (the top activations are 17 MB per
clip per layer; chunking keeps every tensor under 2^31 elements).  Prints one line per batch size."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

CHUNK = 64
sizes = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]
device = torch.device("cuda:0")
torch.cuda.set_device(device)
model, _ = bench.build_model(device)
g = torch.Generator().manual_seed(2)
rows = []
for b in sizes:
    lab = torch.nn.functional.one_hot(torch.arange(b) % 61, 61).float().to(device)
    z = torch.randn(b, 256, generator=g).to(device)

    def run():
        outs = [model.generate_batch(lab[i:i + CHUNK], z[i:i + CHUNK]) for i in range(0, b, CHUNK)]
        return outs[0] if len(outs) == 1 else torch.cat(outs)

    for _ in range(4):          # the third call of a batch size captures its CUDA graph
        out = run()
    reps = 10 if b <= 64 else 3
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        out = run()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    assert tuple(out.shape) == (b, 64000) and bool(torch.isfinite(out).all())
    rows.append(dict(batch=b, ms=round(ms, 3), clips_per_s=round(b / ms * 1e3, 1)))
    print(json.dumps(rows[-1]), flush=True)
    del out
