# round 2, call 4: bias gradient inside the filter-gradient kernel, second-order pixel-norm pair kernel, stage profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_model.log
timeout 600 python tools/tc_stage_profile.py > gpurun_out/stage_profile.txt 2>&1; cat gpurun_out/stage_profile.txt
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; head -40 gpurun_out/step_kernels.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
