set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
python scripts/sweep.py --iters 200 --check --configs 296:0:1:0:-1:2,296:7296:1:0:-1:2,148:14528:1:0:-1:2,296:0:1:0:0.5:2,444:0:1:0:-1:2,296:0:1:512:-1:1 2>&1 | grep "^cfg" | tee gpurun_out/sweep15.log
EHYB_CHUNK=4 python scripts/sweep.py --iters 200 --configs 296:0:1:0:-1:2 2>&1 | grep "^cfg" | sed "s/^/chunk4 /" | tee -a gpurun_out/sweep15.log
EHYB_CHUNK=16 python scripts/sweep.py --iters 200 --configs 296:0:1:0:-1:2 2>&1 | grep "^cfg" | sed "s/^/chunk16 /" | tee -a gpurun_out/sweep15.log
