# end-of-round evidence: full GPU suite, bench line, CUPTI step profile with per-grid convolution detail, ncu launch list of
# the bench command (graph nodes profiled one by one)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --conv-table gpurun_out/conv_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; sed -n 1,8p gpurun_out/step_kernels.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 9000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-spectral --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200; wc -l gpurun_out/launches.csv
python tools/launch_summary.py gpurun_out/launches.csv 2>&1 | head -16
