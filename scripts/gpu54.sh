set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k persistent 2>&1 | tail -4
timeout 500 python scripts/sweep.py --check --configs 592:0:1:0:-1:3,444:0:1:0:-1:3,296:7296:1:0:-1:2 > gpurun_out/sweep20.log 2>&1
grep "^cfg\|rror\|Traceback" gpurun_out/sweep20.log
