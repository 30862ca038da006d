# round 2, call 12: generic spectral kernels, BASELINE config 1 on the CUDA path with media summaries
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spectral_gpu.py -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/pytest_spec.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "config1 or train_and_generate" 2>&1 | tail -12 | tee gpurun_out/pytest_model.log
