set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 10 2>/dev/null | tee gpurun_out/bench_final_n2.json | cut -c1-200
