# round 2 evidence run on the shipped binary: whole GPU suite, bench line (CPU baseline, spectral, per-kernel table), CUPTI
# step profile, ncu launch list of the bench command, `ncu --set full` captures of the dominant kernels, sanitizer passes
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --conv-table gpurun_out/conv_table.txt --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; sed -n 3,12p gpurun_out/step_kernels.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 9000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-spectral --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches.csv; python tools/launch_summary.py gpurun_out/launches.csv 2>&1 | head -14
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:conv_tcw_kernel -s 3 -c 1 -f -o gpurun_out/r2_w_32 python tools/profile_conv.py w 8 128 1024 32 32 1 > gpurun_out/p1.log 2>&1
timeout 300 $NCU -k regex:conv_tck_kernel -s 3 -c 1 -f -o gpurun_out/r2_ck_32 python tools/profile_conv.py c 8 128 1024 32 32 1 > gpurun_out/p2.log 2>&1
timeout 300 $NCU -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/r2_t2_64_32 python tools/profile_conv.py t 8 128 1024 32 64 2 > gpurun_out/p3.log 2>&1
timeout 300 $NCU -k regex:spectrogram_fwd_kernel -s 2 -c 1 -f -o gpurun_out/r2_spec_fwd python tools/profile_spectral.py 256 > gpurun_out/p4.log 2>&1
timeout 300 $NCU -k regex:waveform_fwd_kernel -s 1 -c 1 -f -o gpurun_out/r2_spec_inv python tools/profile_spectral.py 256 > gpurun_out/p5.log 2>&1
tail -1 gpurun_out/p?.log; ls -la gpurun_out/*.ncu-rep
bash tools/gpu_sanitize.sh
