mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29525 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err
tail -c 700 gpurun_out/r2_bench_n2_b.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n2_b.err | tail -20
