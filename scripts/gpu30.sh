set -x
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/topo8.log
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "4-p2p" 2>&1 | tail -4
for cfg in "8 p2p" "8 nccl" "4 p2p"; do
  set -- $cfg
  EHYB_MG_EXCHANGE=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $1 --steps 200 --warmup 10 2> gpurun_out/bench$1_$2_err.log | tee gpurun_out/bench_n$1_$2.json | cut -c1-230
done
