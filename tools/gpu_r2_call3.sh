# round 2, call 3: whole GPU suite on the fused-epilogue binary, step profile, bench with the per-kernel table
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/pytest_gpu_r2.log; tail -25 gpurun_out/pytest_gpu_r2.log
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; head -45 gpurun_out/step_kernels.txt
timeout 600 python bench.py --steps 10 --warmup 3 --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-600 gpurun_out/bench.json; tail -5 gpurun_out/bench.err; cat gpurun_out/kernel_table.txt
