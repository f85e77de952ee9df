set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python scripts/run_rmat.py --scale 22 --blocks --iters 30 2>&1 | tee gpurun_out/rmat22.log | grep -v "^start\|^partition"
timeout 900 python scripts/run_rmat.py --scale 24 --blocks --iters 20 2>&1 | tee gpurun_out/rmat24.log | grep -v "^start\|^partition"
