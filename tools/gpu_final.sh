# end-of-round evidence: ncu launch list of the bench command (graph nodes profiled one by one), CUPTI step profile with
# per-grid convolution detail
mkdir -p gpurun_out
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; sed -n 1,12p gpurun_out/step_kernels.txt; sed -n '/convolution kernels by grid/,$p' gpurun_out/step_kernels.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 9000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-spectral > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300; wc -l gpurun_out/launches.csv
python tools/launch_summary.py gpurun_out/launches.csv 2>&1 | head -30
