set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
python scripts/sweep.py --iters 200 --check --configs 148:0:1:0:0:2:8,148:0:1:0:0:2:4,148:0:1:0:0:2:12,148:0:1:0:0:2:16,148:0:1:0::2:8,148:0:1:256:0:2:8,148:0:1:192:0:2:8,296:0:1:0:0:2:8,296:0:1:0:0:2:4,444:0:1:0:0:2:8 2>&1 | grep "^cfg\|^#" | tee gpurun_out/sweep5.log
