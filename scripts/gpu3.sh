set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
python scripts/sweep.py --iters 200 --check --configs 148:0:1:0::2,148:0:1:1024::1,148:0:2:512::1,148:0:1:128::2,148:0:1:96::2,296:0:1:0::2,296:0:1:256::2,222:0:1:0::2,444:0:1:0::2,148:0:2:0::2 2>&1 | grep "^cfg\|^#" | tee gpurun_out/sweep3.log
