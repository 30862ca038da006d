# round 2 evidence run on the shipped binary: whole GPU suite, bench line (CPU baseline, spectral, per-kernel table), CUPTI
# step profile, ncu launch list of the bench command, `ncu --set full` capture(s), sanitizer passes.  Hard kill timeouts.
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --conv-table gpurun_out/conv_table.txt --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout -s KILL 120 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; sed -n 3,12p gpurun_out/step_kernels.txt
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 9000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-spectral --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches.csv; python tools/launch_summary.py gpurun_out/launches.csv 2>&1 | head -8
NCU="ncu --set full --clock-control none --import-source on"
timeout -s KILL 150 $NCU -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/r2c_t2_64_32 python tools/profile_conv.py t 8 128 1024 32 64 2 > gpurun_out/p3.log 2>&1
timeout -s KILL 150 $NCU -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/r2c_c1_128 python tools/profile_conv.py c 8 32 256 128 128 1 > gpurun_out/p4.log 2>&1
tail -n 1 gpurun_out/p3.log gpurun_out/p4.log; ls -la gpurun_out/r2c_*.ncu-rep
CS="compute-sanitizer --print-limit 30 --error-exitcode 3"
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
for t in memcheck synccheck; do ( time timeout -s KILL 200 $CS --tool $t python -m pytest tests/test_tc_gpu.py -m gpu -q -x ) > gpurun_out/sanitizer_${t}_tc.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_${t}_tc.log | tail -2; done
( time timeout -s KILL 200 $CS --tool memcheck python __graft_entry__.py smoke ) > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; grep -E "ERROR SUMMARY|smoke ok" gpurun_out/sanitizer_memcheck_smoke.log | tail -2
