# full GPU validation: tests, smoke, bench, spectral timing + ncu captures, inference sweep, step profile
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 200 python tools/profile_spectral.py 256 2>&1 | tail -1 | tee gpurun_out/spectral_times.log
timeout 200 python tools/profile_spectral.py 8 2>&1 | tail -1 | tee -a gpurun_out/spectral_times.log
timeout 400 python bench.py --steps 10 --warmup 3 --conv-table gpurun_out/conv_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python tools/inference_sweep.py > gpurun_out/inference_sweep.txt 2>&1; cat gpurun_out/inference_sweep.txt | tail -12
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:spectrogram_fwd_kernel -s 2 -c 1 -f -o gpurun_out/spec_fwd python tools/profile_spectral.py 256 > gpurun_out/ps1.log 2>&1
timeout 300 $NCU -k regex:waveform_fwd_kernel -s 1 -c 1 -f -o gpurun_out/spec_inv python tools/profile_spectral.py 256 > gpurun_out/ps2.log 2>&1
timeout 300 python tools/step_profile.py 3 > gpurun_out/step_kernels.txt 2>&1; head -30 gpurun_out/step_kernels.txt
