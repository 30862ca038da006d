# round 2, call 10: tcw converter-warp experiment (8 vs 16), bench with the per-kernel table
mkdir -p gpurun_out
W="w 8 128 1024 32 32 1 bias"; W2="w 8 64 512 64 64 1 bias"; W3="w 8 128 1024 32 64 2 bias"; W4="w 8 16 128 256 256 1 bias"
for lib in prof prof16; do for c in "$W" "$W2" "$W3" "$W4"; do echo "== GS_LIB=$lib"; GS_LIB=$lib timeout 120 python tools/tc_stage_profile.py $c; done; done > gpurun_out/stage_profile_tcw.txt 2>&1
echo "== GS_LIB=prof16 GS_TCW_NO_NSTACK=1" >> gpurun_out/stage_profile_tcw.txt; GS_LIB=prof16 GS_TCW_NO_NSTACK=1 timeout 120 python tools/tc_stage_profile.py $W >> gpurun_out/stage_profile_tcw.txt 2>&1
cat gpurun_out/stage_profile_tcw.txt
GS_LIB=prof16 timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -k "filter_gradient" 2>&1 | tail -3 | tee gpurun_out/pytest_tcw16.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-spectral --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json; tail -5 gpurun_out/bench.err; cat gpurun_out/kernel_table.txt
