"""Which ATen operators (not this repo's kernels) run in one EAGER training iteration, with input shapes and the stack of
the Python frame that called them.  python tools/aten_ops_profile.py > profiles/aten_ops.txt"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

device = torch.device("cuda", 0)
model, store = bench.build_model(device)
model.use_cuda_graphs = False
host = bench.host_batches(3, 0, pinned=True)
dev = [[t.to(device) for t in b] for b in host]
bench.run_steps(model, dev[:2], device, False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    bench.run_steps(model, dev[2:], device, False)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0, None])
for ev in prof.events():
    if not ev.name.startswith("aten::") or ev.device_type != torch.autograd.DeviceType.CPU:
        continue
    us = getattr(ev, "device_time_total", 0.0) or getattr(ev, "cuda_time_total", 0.0)
    if us <= 0 or ev.name in ("aten::empty", "aten::empty_like", "aten::view", "aten::reshape", "aten::as_strided", "aten::slice"):
        continue
    frame = next((s for s in (ev.stack or []) if "gansynth_b200" in s or "bench.py" in s), "?")
    key = (ev.name, str(ev.input_shapes)[:70], frame[-60:])
    agg[key][0] += 1
    agg[key][1] += us
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%8.1f us n=%3d  %-22s %-70s %s" % (v[1], v[0], k[0], k[1], k[2]))
