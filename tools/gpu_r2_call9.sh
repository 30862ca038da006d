# round 2, call 9: gradient sinks, persistent split-weight cache with refresh, dense kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_tc.log
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_main_gpu.py tests/test_dataset_gpu.py tests/test_spectral_gpu.py -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/pytest_model.log
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; head -50 gpurun_out/step_kernels.txt | cut -c1-100
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-spectral --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
