set -x
nvidia-smi -L | wc -l
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 200 --warmup 10 2>gpurun_out/bench${N}_err.log | tee gpurun_out/bench_n$N.json | cut -c1-700
done
python bench.py --gpus 1 --steps 200 --warmup 10 2>/dev/null | tee gpurun_out/bench_n1_c.json | cut -c1-300
