python scripts/ref_gpu_compare.py --kind st27 --dims 128 128 128 2>&1 | tail -3 | tee gpurun_out/ref_gpu_c2.json
python scripts/ref_gpu_compare.py --kind lap2d --dims 1024 1024 --iters 500 2>&1 | tail -1 | tee gpurun_out/ref_gpu_c1.json
