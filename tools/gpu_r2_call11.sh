# round 2, call 11: pitch-classifier kernels / network, general kernel sizes, refresh test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tc_gpu.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "resnet or small_forward" 2>&1 | tail -8 | tee gpurun_out/pytest_model.log
