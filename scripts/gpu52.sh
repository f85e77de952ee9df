set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "power_law or accuracy_gate" 2>&1 | tail -4
timeout 400 python scripts/run_rmat.py --scale 22 --blocks --iters 30 2>&1 | tee gpurun_out/rmat22_staged.log | grep "gate\|us per product\|cuSPARSE"
EHYB_OVF_STAGED_MIN=2000000000 timeout 400 python scripts/run_rmat.py --scale 22 --blocks --iters 30 2>&1 | grep "us per product"
