set -x
for d in 0 4; do EHYB_DEBUG_SKIP=$d python scripts/sweep.py --iters 200 --configs 148:14528:1:0:-1:2,148:14528:1:224:-1:2,296:0:1:0:-1:2 2>&1 | grep "^cfg" | sed "s/^/dbg=$d /"; done | tee gpurun_out/sweep14.log
