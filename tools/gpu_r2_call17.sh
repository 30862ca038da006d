# round 2, call 17: two-group epilogue in conv_tc, second attempt -- hard kill timeouts
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests/test_tc_gpu.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_tc.log
