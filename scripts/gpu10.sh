set -x
nvidia-smi -L
python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 2>gpurun_out/bench2_err.log | tee gpurun_out/bench_n2.json
tail -20 gpurun_out/bench2_err.log
