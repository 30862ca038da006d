mkdir -p gpurun_out
timeout -s KILL 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-spectral > gpurun_out/bench_4gpu_fused.json 2> gpurun_out/bench_4gpu_fused.err
cut -c1-260 gpurun_out/bench_4gpu_fused.json; grep -E "Error|error" gpurun_out/bench_4gpu_fused.err | tail -3
