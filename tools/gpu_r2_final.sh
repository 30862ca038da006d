# round 2, last call: whole GPU suite and smoke() on the final tree -- hard kill timeouts
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
