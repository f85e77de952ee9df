mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_t1.log 2>&1; tail -15 gpurun_out/r2_t1.log
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; tail -c 1500 gpurun_out/r2_bench_n1_a.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2_bench_n2_a.json 2> gpurun_out/r2_bench_n2_a.err; tail -c 1500 gpurun_out/r2_bench_n2_a.json; tail -5 gpurun_out/r2_bench_n2_a.err
EHYB_MG_PLAN=staged timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2_bench_n2_staged.json 2> gpurun_out/r2_bench_n2_staged.err; tail -c 600 gpurun_out/r2_bench_n2_staged.json
