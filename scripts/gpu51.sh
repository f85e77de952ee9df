set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py 2>/dev/null | tee gpurun_out/bench_head_n1.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | cut -c1-200
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 10 2>/dev/null | tee gpurun_out/bench_head_n2.json | cut -c1-200
