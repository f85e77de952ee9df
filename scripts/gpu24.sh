set -x
nvidia-smi topo -m 2>&1 | head -8
timeout 200 python bench.py --steps 100 --warmup 5 2>gpurun_out/b1_err.log | tee gpurun_out/bench_n1_p2pbuild.json | cut -c1-260
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -25
for ex in p2p nccl; do
  EHYB_MG_EXCHANGE=$ex timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 2> gpurun_out/bench2_${ex}_err.log | tee gpurun_out/bench_n2_${ex}.json | cut -c1-500
  tail -5 gpurun_out/bench2_${ex}_err.log
done
