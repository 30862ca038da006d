# round 2, call 8: shared per-device context, vectorised dense kernels, epilogue phase timers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_model.log
timeout 120 python tools/tc_stage_profile.py c 8 128 1024 32 32 1 > gpurun_out/stage_profile_tck.txt 2>&1; cat gpurun_out/stage_profile_tck.txt
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; head -45 gpurun_out/step_kernels.txt | cut -c1-100
