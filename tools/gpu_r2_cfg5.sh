#!/bin/bash
# short bench run without the CPU legs: checks the config-5 roofline fields on the GPU
mkdir -p gpurun_out
timeout -s KILL 200 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
echo rc=$?
python - <<'P'
import json
l = json.loads(open("gpurun_out/bench_cfg5.json").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["e2e"]["value"])
print(json.dumps(l["secondary"]["inference"], indent=1))
P
tail -3 gpurun_out/bench_cfg5.err
