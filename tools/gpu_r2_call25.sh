mkdir -p gpurun_out
timeout -s KILL 100 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "max_pool or group_norm or momentum" 2>&1 | grep -E "passed|failed|^E |Error" | head -8
timeout -s KILL 150 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "pitch_classifier or resnet" 2>&1 | grep -E "passed|failed|^E |Error" | head -12
