mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_gpu_final2.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu_final2.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
