mkdir -p gpurun_out
timeout 240 python scripts/rmat_variants.py --scale 24 --iters 30 > gpurun_out/r2_rmat24.log 2>&1; cat gpurun_out/r2_rmat24.log | tail -12
