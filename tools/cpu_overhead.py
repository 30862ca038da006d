"""Host-side launch cost of one training iteration: wall time to ENQUEUE a step (no synchronisation inside)
against the device time of the same steps.  python tools/cpu_overhead.py [steps]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
device = torch.device("cuda", 0)
model, store = bench.build_model(device)
host = bench.host_batches(steps + 3, 0, pinned=True)
dev = [[t.to(device) for t in b] for b in host]
bench.run_steps(model, dev[:3], device, False)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
s.record()
bench.run_steps(model, dev[3:], device, False)
e.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("enqueue %.2f ms/step (host), device %.2f ms/step, wall %.2f ms/step" % (
    (t1 - t0) * 1e3 / steps, s.elapsed_time(e) / steps, (t2 - t0) * 1e3 / steps))
