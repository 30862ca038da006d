# round 2, call 5: tck epilogue split across both warpgroups (channel halves), early accumulator release
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_model.log
for c in "c 8 128 1024 32 32 1" "t 8 128 1024 32 32 1 mask" "c 8 128 1024 32 32 1 pn" "c 8 64 512 64 64 1" "t 8 64 512 64 64 1 mask"; do timeout 120 python tools/tc_stage_profile.py $c; done > gpurun_out/stage_profile_tck.txt 2>&1; cat gpurun_out/stage_profile_tck.txt
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; head -30 gpurun_out/step_kernels.txt | cut -c1-100
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-spectral --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
