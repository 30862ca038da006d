#!/bin/bash
# one short call: only the reference-vector GPU tests (budget is nearly spent)
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_zz_reference_vectors_gpu.py -q -m gpu -x -s ${GS_K:+-k "$GS_K"} 2>&1 | grep -v Warning | tail -40 > gpurun_out/refvec_gpu.txt
cat gpurun_out/refvec_gpu.txt
