set -x
timeout 600 python -m pytest tests/test_gpu_solver.py -x -q 2>&1 | tail -12
timeout 400 python scripts/pcg_bench.py 2>/dev/null | tee gpurun_out/pcg_bench.log
