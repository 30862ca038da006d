# round 2, call 18: model tests and step profile on the two-group conv_tc epilogue -- hard kill timeouts
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_model.log
timeout -s KILL 100 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; sed -n 3,36p gpurun_out/step_kernels.txt | cut -c1-100
