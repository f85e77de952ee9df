set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
( EHYB_DYNAMIC_DEAL=0 timeout 200 python scripts/mg_trace.py n1_static
  EHYB_DYNAMIC_DEAL=1 timeout 200 python scripts/mg_trace.py n1_dynamic
  EHYB_DYNAMIC_DEAL=1 EHYB_DEBUG_SKIP=3 timeout 200 python scripts/mg_trace.py n1_dynamic_nomath
) 2>gpurun_out/trace_err.log | grep -v "^\*\*\*\|OMP_NUM" | tee gpurun_out/trace_summary_f.log
tail -3 gpurun_out/trace_err.log
