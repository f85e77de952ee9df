set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516"
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 10 2>gpurun_out/bench_head_n2_err.log | tee gpurun_out/bench_head_n2.json | cut -c1-200
tail -3 gpurun_out/bench_head_n2_err.log | cut -c1-200
timeout 200 $TR bench.py --impl reference --gpus 2 --steps 5 --warmup 3 2>/dev/null | cut -c1-160
EHYB_MG_EXCHANGE=nccl timeout 300 $TR bench.py --gpus 2 --steps 50 --warmup 5 2>/dev/null | cut -c1-200
