# round 2, call 2: fused epilogues (mask / pixel norm) -- kernel tests, model parity, step profile, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/pytest_model.log
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; head -45 gpurun_out/step_kernels.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
