set -x
for c in "8 4" "4 4" "8 8" "16 8"; do set -- $c; EHYB_CHUNK=$1 EHYB_CHUNK_REM=$2 python scripts/sweep.py --iters 200 --check --configs 148:14528:1:0:-1:2 2>&1 | grep "^cfg" | sed "s/^/chunk=$1,$2 /"; done | tee gpurun_out/sweep13.log
