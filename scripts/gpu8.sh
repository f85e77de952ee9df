set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for kr in 4 8 12; do EHYB_CHUNK_REM=$kr python scripts/sweep.py --iters 200 --check --configs 148:0:1:0:0:2:8,148:0:1:0::2:8,148:0:1:0:0.25:2:8 2>&1 | grep "^cfg" | sed "s/^/kr=$kr /"; done | tee gpurun_out/sweep8.log
