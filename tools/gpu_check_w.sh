mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -x -q -k "filter_gradient" 2>&1 | tail -12 > gpurun_out/w_tests.log
cat gpurun_out/w_tests.log
SHAPES=("w 8 128 1024 32 32 1" "w 8 64 512 64 64 1" "w 8 32 256 128 128 1" "w 8 16 128 256 256 1" "w 8 8 64 256 256 1" "w 8 2 16 256 256 1" "w 8 128 1024 32 64 2" "w 8 64 512 64 128 2" "w 8 32 256 128 256 2" "w 8 16 128 256 256 2")
for a in "${SHAPES[@]}"; do timeout 120 python tools/profile_conv.py $a 2>&1 | tail -1; done > gpurun_out/w_times.log 2>&1
cat gpurun_out/w_times.log
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_tcw_kernel -s 3 -c 1 -f -o gpurun_out/w_32_v2 python tools/profile_conv.py w 8 128 1024 32 32 1 > gpurun_out/p2.log 2>&1
$NCU -k regex:conv_tcw_kernel -s 3 -c 1 -f -o gpurun_out/w_256_v2 python tools/profile_conv.py w 8 16 128 256 256 1 > gpurun_out/p4.log 2>&1
