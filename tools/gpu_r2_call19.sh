# round 2, call 19: 1x1 reduce / filter-gradient kernels -- hard kill timeouts
mkdir -p gpurun_out
timeout -s KILL 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_k.log
timeout -s KILL 100 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; grep -E "kernel time|conv1x1|col_sum|pixel_norm_vec_kernel<2, 1" gpurun_out/step_kernels.txt | cut -c1-100
