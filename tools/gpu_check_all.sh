mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --conv-table gpurun_out/conv_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; head -45 gpurun_out/step_kernels.txt
