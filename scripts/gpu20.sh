set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
python scripts/sweep.py --iters 200 --check --configs 296:0:1:0:0.5:2,296:0:1:0:-1:2 2>&1 | grep "^cfg"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 2>gpurun_out/bench2_err.log | tee gpurun_out/bench_n2_b.json | cut -c1-330
