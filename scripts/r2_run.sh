mkdir -p gpurun_out
timeout 600 python scripts/rmat_variants.py --scale 24 --iters 30 --variants hubs1k,hubs2k,hubs4k,hubs6k,hubs8k,hubs12k > gpurun_out/r2_rmat24_v8_hubs.log 2>&1; tail -9 gpurun_out/r2_rmat24_v8_hubs.log
