set -x
EHYB_BENCH_GRID=256x256x256 EHYB_BENCH_SCALING=strong timeout 1200 python bench.py --steps 50 --warmup 5 2>gpurun_out/b256_err.log | tee gpurun_out/bench_strong256_n1.json | cut -c1-400
tail -3 gpurun_out/b256_err.log
