mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_t2.log 2>&1; tail -15 gpurun_out/r2_t2.log
EHYB_BENCH_GRID=256x256x256 timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_grid256_n1.json 2> gpurun_out/r2_grid256_n1.err; tail -c 2500 gpurun_out/r2_grid256_n1.json; tail -5 gpurun_out/r2_grid256_n1.err
EHYB_BENCH_GRID=256x256x256 timeout 600 $TR --master-port 29531 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_grid256_n2.json 2> gpurun_out/r2_grid256_n2.err; tail -c 2500 gpurun_out/r2_grid256_n2.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_grid256_n2.err | tail -5
timeout 900 $TR --master-port 29532 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_grid512_n2.json 2> gpurun_out/r2_grid512_n2.err; tail -c 2500 gpurun_out/r2_grid512_n2.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_grid512_n2.err | tail -5
