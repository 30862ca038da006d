mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_kernels_gpu.py -x -q 2>&1 | tail -6 > gpurun_out/w_tests.log
cat gpurun_out/w_tests.log
SHAPES=("w 8 128 1024 32 32 1" "w 8 64 512 64 64 1" "w 8 32 256 128 128 1" "w 8 16 128 256 256 1" "w 8 8 64 256 256 1" "w 8 128 1024 32 64 2" "w 8 32 256 128 256 2")
for cfg in "A=1"; do
  echo "== cfg: $cfg"
  for a in "${SHAPES[@]}"; do env $cfg timeout 120 python tools/profile_conv.py $a 2>&1 | tail -1; done
done > gpurun_out/w_times.log 2>&1
cat gpurun_out/w_times.log
