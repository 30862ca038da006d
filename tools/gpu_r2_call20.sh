mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-spectral --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_q.json"))
print(d["ms_per_step"], d["value"], d["config"].get("eager_ms_per_step"), d["config"].get("eager_host_enqueue_ms_per_step"))
r=d["roofline"]; print(r["kernel"], r["bound"], round(r["frac"],3), round(r["avg_launch_ms"]*1e3,1), r["function_share_of_step"], r["share_of_eager_step"])
PY
tail -3 gpurun_out/bench_q.err
