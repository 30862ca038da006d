mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/tc_tests.log
cat gpurun_out/tc_tests.log
for a in "c 8 128 1024 32 32 1" "t 8 128 1024 32 32 1" "c 8 64 512 64 64 1" "c 8 32 256 128 128 1" "c 8 16 128 256 256 1" "c 8 4 32 256 256 1" "c 8 2 16 256 256 1" "c 8 128 1024 32 64 2" "t 8 128 1024 32 64 2" "c 8 64 512 64 128 2" "t 8 64 512 64 128 2" "c 8 8 64 256 256 2" "t 8 8 64 256 256 2" "w 8 128 1024 32 32 1" "w 8 64 512 64 64 1" "w 8 8 64 256 256 1" "w 8 128 1024 32 64 2"; do timeout 120 python tools/profile_conv.py $a; done > gpurun_out/conv_times.log 2>&1
cat gpurun_out/conv_times.log
