set -x
python scripts/ref_gpu_compare.py --kind st27 --dims 128 128 128 2>&1 | grep "^{" | tee gpurun_out/ref_gpu_c2.json
python scripts/ref_gpu_compare.py --kind lap2d --dims 1024 1024 --iters 500 2>&1 | grep "^{" | tee gpurun_out/ref_gpu_c1.json
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
python scripts/sweep.py --iters 200 --check --configs 296:0:1:0:0.5:2 2>&1 | grep "^cfg"
