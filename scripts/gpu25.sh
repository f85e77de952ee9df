set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -8
for ex in p2p; do
  EHYB_MG_EXCHANGE=$ex timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 2> gpurun_out/bench2_${ex}_err.log | tee gpurun_out/bench_n2_${ex}_b.json | cut -c1-230
done
# N=1: warps per CTA (static deal of ~111 slices per partition over the warps)
CUDA_VISIBLE_DEVICES=0 timeout 400 python scripts/sweep.py --check --configs 296:7296:1:768:-1,296:7296:1:736:-1,296:7296:1:704:-1,296:7296:1:608:-1,296:7296:1:640:-1 > gpurun_out/sweep16.log 2>&1
cat gpurun_out/sweep16.log
