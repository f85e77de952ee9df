set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
( CUDA_VISIBLE_DEVICES=0 timeout 200 python scripts/mg_trace.py n1c
  timeout 300 $TR scripts/mg_trace.py n2c_order1
) 2>gpurun_out/trace_err.log | grep -v "^\*\*\*\|OMP_NUM" | tee gpurun_out/trace_summary_c.log
tail -3 gpurun_out/trace_err.log
CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --steps 200 --warmup 10 2>gpurun_out/b1_err.log | tee gpurun_out/bench_n1_d.json | cut -c1-230
EHYB_MG_EXCHANGE=p2p timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 10 2> gpurun_out/bench2_p2p_err.log | tee gpurun_out/bench_n2_p2p_d.json | cut -c1-230
