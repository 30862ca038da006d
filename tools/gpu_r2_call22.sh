mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "pitch_classifier" 2>&1 | grep -E "passed|failed|^E |Error" | head -12 | tee gpurun_out/pytest_m.log
