set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
( CUDA_VISIBLE_DEVICES=0 timeout 200 python scripts/mg_trace.py n1
  timeout 300 $TR scripts/mg_trace.py n2_order1
  EHYB_P2P_ORDER=0 timeout 300 $TR scripts/mg_trace.py n2_order0
  EHYB_P2P_ORDER=0 EHYB_P2P_PUSH_CTAS=32 timeout 300 $TR scripts/mg_trace.py n2_order0_push32
  EHYB_P2P_PUSH_CTAS=16 timeout 300 $TR scripts/mg_trace.py n2_order1_push16
) 2>gpurun_out/trace_err.log | grep -v "^\*\*\*\|OMP_NUM" | tee gpurun_out/trace_summary.log
tail -5 gpurun_out/trace_err.log
