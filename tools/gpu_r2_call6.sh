# round 2, call 6: caller-owned workspace (gs_context), tck without "cat" at 32 channels + 4 accumulator buffers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_kernels_gpu.py tests/test_spectral_gpu.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_main_gpu.py tests/test_dataset_gpu.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_model.log
for cat in 0 1; do for c in "c 8 128 1024 32 32 1" "t 8 128 1024 32 32 1 mask" "c 8 128 1024 32 32 1 pn"; do echo "GS_TCK_CAT=$cat"; GS_TCK_CAT=$cat timeout 120 python tools/tc_stage_profile.py $c; done; done > gpurun_out/stage_profile_tck.txt 2>&1; cat gpurun_out/stage_profile_tck.txt
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; head -30 gpurun_out/step_kernels.txt | cut -c1-100
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-spectral --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
