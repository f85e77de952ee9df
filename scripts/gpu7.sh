set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for rw in 0 4 8 12; do EHYB_REM_WARPS=$rw python scripts/sweep.py --iters 200 --check --configs 148:0:1:0:0:2:8,148:0:1:0::2:8 2>&1 | grep "^cfg" | sed "s/^/remwarps=$rw /"; done | tee gpurun_out/sweep7.log
EHYB_REM_WARPS=8 EHYB_DEBUG_SKIP=2 python scripts/sweep.py --iters 200 --configs 148:0:1:0:0:2:8 2>&1 | grep "^cfg" | sed "s/^/skipELL /" | tee -a gpurun_out/sweep7.log
