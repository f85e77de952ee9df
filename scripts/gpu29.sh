set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4
( timeout 300 $TR scripts/mg_trace.py n2d_order1
  EHYB_P2P_ORDER=0 timeout 300 $TR scripts/mg_trace.py n2d_order0
) 2>gpurun_out/trace_err.log | grep -v "^\*\*\*\|OMP_NUM" | tee gpurun_out/trace_summary_d.log
tail -3 gpurun_out/trace_err.log
EHYB_MG_EXCHANGE=p2p timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 10 2> gpurun_out/bench2_p2p_err.log | tee gpurun_out/bench_n2_p2p_e.json | cut -c1-230
CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --steps 200 --warmup 10 2>gpurun_out/b1_err.log | tee gpurun_out/bench_n1_e.json | cut -c1-230
