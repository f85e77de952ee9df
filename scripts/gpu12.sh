set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
python scripts/sweep.py --iters 200 --check --configs 148:14528:1:0:-1:2,148:14528:1:0:0.5:2,148:14528:1:256:-1:2,296:0:1:0:-1:2 2>&1 | grep "^cfg" | tee gpurun_out/sweep12.log
for d in 1 2 3; do EHYB_DEBUG_SKIP=$d python scripts/sweep.py --iters 200 --configs 148:14528:1:0:-1:2 2>&1 | grep "^cfg" | sed "s/^/skip=$d /"; done | tee -a gpurun_out/sweep12.log
