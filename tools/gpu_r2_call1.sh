# round 2, call 1: baseline of the round-1 binary on this box + sanitizer evidence + tanh flake hunt
mkdir -p gpurun_out
timeout 300 python tools/flake_hunt.py 600 2>&1 | tail -8 | tee gpurun_out/flake_hunt.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_base.json 2> gpurun_out/bench_r2_base.err
cut -c1-600 gpurun_out/bench_r2_base.json
bash tools/gpu_sanitize.sh
