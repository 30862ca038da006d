"""Kernel-time breakdown of graph-replayed training iterations with torch.profiler (CUPTI activity records, no
replay overhead).  python tools/step_profile.py [steps] > profiles/step_kernels.txt"""
import collections
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
device = torch.device("cuda", 0)
model, store = bench.build_model(device)
host = bench.host_batches(steps + 4, 0, pinned=True)
dev = [[t.to(device) for t in b] for b in host]
bench.run_steps(model, dev[:4], device, False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    bench.run_steps(model, dev[4:], device, False)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name = re.sub(r"\(.*", "", ev.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))[:70]
    us = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    agg[name][0] += 1
    agg[name][1] += us
    tot += us
print("kernel time %.2f ms/step over %d steps, %d launches/step" % (tot / 1e3 / steps, steps, sum(v[0] for v in agg.values()) // steps))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%9.1f us/step %5.1f%% n=%4d avg=%7.1f  %s" % (v[1] / steps, 100 * v[1] / tot, v[0] // steps, v[1] / v[0], k))

# per-(kernel, grid) detail of the convolution family from the chrome trace (kineto records grid / block there)
import json  # noqa: E402
import tempfile  # noqa: E402

path = os.path.join(tempfile.gettempdir(), "gs_step_trace.json")
prof.export_chrome_trace(path)
detail = collections.defaultdict(lambda: [0, 0.0, []])
for ev in json.load(open(path)).get("traceEvents", []):
    args = ev.get("args") or {}
    if ev.get("cat") != "kernel" or "grid" not in args or "conv_" not in ev.get("name", ""):
        continue
    name = re.sub(r"\(.*", "", ev["name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", ""))[:40]
    key = (name, tuple(args["grid"]))
    detail[key][0] += 1
    detail[key][1] += ev.get("dur", 0.0)
    detail[key][2].append(ev.get("dur", 0.0))
print("\nconvolution kernels by grid (us per launch: mean [min .. max], launches per step, us per step):")
for (name, grid), v in sorted(detail.items(), key=lambda kv: -kv[1][1])[:60]:
    print("%8.1f [%6.1f .. %6.1f]  n=%3d  %8.1f  %-40s grid %s" % (v[1] / v[0], min(v[2]), max(v[2]), v[0] // steps, v[1] / steps,
                                                            name, "x".join(str(g) for g in grid)))
