# compute-sanitizer evidence: memcheck / racecheck / initcheck over the kernel-level GPU tests and one
# small training iteration.  Logs -> gpurun_out/sanitizer_*.log (summaries are copied to profiles/ by hand).
mkdir -p gpurun_out
CS="compute-sanitizer --print-limit 30 --error-exitcode 3"
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
run() {  # name, timeout, tool args..., -- command
  name=$1; shift; tmo=$1; shift
  ( time timeout $tmo "$@" ) > gpurun_out/sanitizer_$name.log 2>&1
  echo "== $name: exit $? ==" | tee -a gpurun_out/sanitizer_$name.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|real" gpurun_out/sanitizer_$name.log | tail -6
}
run memcheck_kernels 600 $CS --tool memcheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x
run memcheck_tc 600 $CS --tool memcheck python -m pytest tests/test_tc_gpu.py -m gpu -q -x
run memcheck_smoke 420 $CS --tool memcheck python __graft_entry__.py smoke
GS_CONV_TC=0 run initcheck_kernels_fp32 300 $CS --show-backtrace no --tool initcheck python -m pytest tests/test_kernels_gpu.py -m gpu -q
run racecheck_kernels 600 $CS --tool racecheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x
run synccheck_tc 600 $CS --tool synccheck python -m pytest tests/test_tc_gpu.py -m gpu -q -x
unset PYTORCH_NO_CUDA_MEMORY_CACHING
