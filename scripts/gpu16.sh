python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -12
EHYB_CHUNK=4 python scripts/sweep.py --iters 200 --configs 296:0:1:0:-1:2,296:0:1:576:-1:2,296:0:1:512:-1:2,296:0:1:448:-1:2 2>&1 | grep "^cfg" | sed "s/^/chunk4 /"
EHYB_CHUNK=4 EHYB_DEBUG_SKIP=3 python scripts/sweep.py --iters 200 --configs 296:0:1:0:-1:2 2>&1 | grep "^cfg" | sed "s/^/chunk4 skip3 /"
EHYB_CHUNK=8 EHYB_DEBUG_SKIP=3 python scripts/sweep.py --iters 200 --configs 296:0:1:0:-1:2 2>&1 | grep "^cfg" | sed "s/^/chunk8 skip3 /"
