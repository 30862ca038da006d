# round 2, call 14b (2 GPUs): fused all-reduce + Adam over NVLink symmetric memory -- correctness, then bench with / without
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/tools/fused_allreduce_check.py 2>&1 | grep -E "iteration|CHECK|Error|error" | tee gpurun_out/fused_check.log
for mode in 0 1 p2p; do
GS_FUSED_ALLREDUCE=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-spectral > gpurun_out/bench_2gpu_fused_$mode.json 2> gpurun_out/bench_2gpu_fused_$mode.err
echo "mode $mode: $(cut -c1-220 gpurun_out/bench_2gpu_fused_$mode.json)"; grep -E "Error|error" gpurun_out/bench_2gpu_fused_$mode.err | tail -3
done
