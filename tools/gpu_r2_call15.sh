# round 2, call 15 (8 GPUs): fused all-reduce + Adam at world size 8 -- check, then bench fused (multimem) vs NCCL
mkdir -p gpurun_out
N=8
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/tools/fused_allreduce_check.py 2>&1 | grep -E "iteration|update only|CHECK|Error|error" | tee gpurun_out/fused_check_8gpu.log
for mode in 1 0; do
GS_FUSED_ALLREDUCE=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-spectral > gpurun_out/bench_8gpu_fused_$mode.json 2> gpurun_out/bench_8gpu_fused_$mode.err
echo "mode $mode: $(cut -c1-220 gpurun_out/bench_8gpu_fused_$mode.json)"; grep -E "Error|error" gpurun_out/bench_8gpu_fused_$mode.err | tail -3
done
