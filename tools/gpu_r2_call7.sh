# round 2, call 7: epilogue phase timers; the tc-0.3 step-parity failure in detail
mkdir -p gpurun_out
for c in "c 8 128 1024 32 32 1" "c 8 64 512 64 64 1"; do timeout 120 python tools/tc_stage_profile.py $c; done > gpurun_out/stage_profile_tck.txt 2>&1; cat gpurun_out/stage_profile_tck.txt
for i in 1 2 3; do timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "test_small_step_parity" 2>&1 | grep -E "passed|failed|Error|beyond|assert " | head -8; done | tee gpurun_out/pytest_small_step.log
