set -x
( EHYB_L2_HINT=0 timeout 200 python scripts/mg_trace.py n1_hint0
  EHYB_L2_HINT=1 timeout 200 python scripts/mg_trace.py n1_hint1
  EHYB_L2_HINT=1 EHYB_WIN_PIECE=4096 timeout 200 python scripts/mg_trace.py n1_hint1_piece4k
  EHYB_L2_HINT=0 EHYB_WIN_PIECE=4096 timeout 200 python scripts/mg_trace.py n1_hint0_piece4k
) 2>gpurun_out/trace_err.log | grep -v "^\*\*\*\|OMP_NUM" | tee gpurun_out/trace_summary_e.log
tail -3 gpurun_out/trace_err.log
