set -x
for d in 0 1 2 3; do EHYB_DEBUG_SKIP=$d python scripts/sweep.py --iters 200 --configs 148:0:1:0:0:2:8,148:0:1:0::2:8 2>&1 | grep "^cfg" | sed "s/^/skip=$d /"; done | tee gpurun_out/sweep6.log
python scripts/sweep.py --dims 192 192 192 --iters 100 --check --configs 444:0:1:0:0:2:8,296:0:1:0:0:2:8 2>&1 | grep "^cfg\|^#" | tee -a gpurun_out/sweep6.log
