set -x
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
EHYB_BENCH_GRID=256x256x256 EHYB_BENCH_SCALING=strong timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 10 2> gpurun_out/bench_strong256_n${N}_err.log | tee gpurun_out/bench_strong256_n$N.json | cut -c1-260
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 2> gpurun_out/bench_weak_n${N}_err.log | tee gpurun_out/bench_weak_n$N.json | cut -c1-260
if [ "$N" = "2" ]; then
  EHYB_MG_EXCHANGE=nccl timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 2> gpurun_out/bench_weak_nccl_n${N}_err.log | tee gpurun_out/bench_weak_nccl_n$N.json | cut -c1-260
  timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
fi
