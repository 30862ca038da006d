# training-step validation: GPU tests, smoke, bench, CUPTI step profile
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 10 --warmup 3 --conv-table gpurun_out/conv_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; head -24 gpurun_out/step_kernels.txt
