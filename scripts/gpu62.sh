timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -k "oneway" 2>&1 | tail -12
