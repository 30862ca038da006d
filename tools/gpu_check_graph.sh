mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -k "graph or entry_points" 2>&1 | tail -15 > gpurun_out/graph_tests.log
cat gpurun_out/graph_tests.log
timeout 300 python tools/cpu_overhead.py 5 2>&1 | tail -3
GS_CUDA_GRAPHS=0 timeout 300 python tools/cpu_overhead.py 5 2>&1 | tail -3
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spectral > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; cat gpurun_out/bench_graph.json; tail -5 gpurun_out/bench_graph.err
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/c1_256s_v2 python tools/profile_conv.py c 8 4 32 256 256 1 > gpurun_out/p3.log 2>&1
