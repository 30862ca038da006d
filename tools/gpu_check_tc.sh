mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/tc_tests.log
cat gpurun_out/tc_tests.log
timeout 120 python tools/mma_timing.py > gpurun_out/mma_timing.log 2>&1
cat gpurun_out/mma_timing.log
