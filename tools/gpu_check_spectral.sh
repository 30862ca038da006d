# spectral kernels: parity tests, batch-256 timing, one ncu --set full capture of each kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spectral_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_spectral.log
cat gpurun_out/pytest_spectral.log
timeout 200 python tools/profile_spectral.py 256 2>&1 | tail -3 | tee gpurun_out/spectral_times.log
timeout 200 python tools/profile_spectral.py 8 2>&1 | tail -1 | tee -a gpurun_out/spectral_times.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:spectrogram_fwd_kernel -s 2 -c 1 -f -o gpurun_out/spec_fwd python tools/profile_spectral.py 256 > gpurun_out/ps1.log 2>&1
timeout 400 $NCU -k regex:waveform_fwd_kernel -s 1 -c 1 -f -o gpurun_out/spec_inv python tools/profile_spectral.py 256 > gpurun_out/ps2.log 2>&1
tail -n 2 gpurun_out/ps1.log; tail -n 2 gpurun_out/ps2.log
