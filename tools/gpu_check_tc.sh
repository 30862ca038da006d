mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -x -q 2>&1 | tail -12 > gpurun_out/tc_tests.log
cat gpurun_out/tc_tests.log
SHAPES=("c 8 128 1024 32 32 1" "t 8 128 1024 32 32 1" "c 8 64 512 64 64 1")
for cfg in "A=1" "GS_TC_NO_KSTACK=1"; do
  echo "== cfg: $cfg"
  for a in "${SHAPES[@]}"; do env $cfg timeout 120 python tools/profile_conv.py $a 2>&1 | tail -1; done
done > gpurun_out/conv_times.log 2>&1
cat gpurun_out/conv_times.log
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_tck_kernel -s 3 -c 1 -f -o gpurun_out/ck_32 python tools/profile_conv.py c 8 128 1024 32 32 1 > gpurun_out/p1.log 2>&1
