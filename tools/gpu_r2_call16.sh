# round 2, call 16: two-group epilogue in conv_tc (all forms), pixel-norm kernel without 64-bit divisions
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_model.log
for c in "t 8 128 1024 32 64 2" "c 8 128 1024 32 64 2" "c 8 32 256 128 128 1"; do timeout 100 python tools/tc_stage_profile.py $c; done > gpurun_out/stage_profile_tc2.txt 2>&1; cat gpurun_out/stage_profile_tc2.txt
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; sed -n 3,40p gpurun_out/step_kernels.txt | cut -c1-100
