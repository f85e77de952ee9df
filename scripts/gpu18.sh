set -x
python scripts/run_rmat.py --scale 20 2>&1 | grep -v "^nParts\|k-way\|partition fin" | tee gpurun_out/rmat20.log
python scripts/run_rmat.py --scale 22 --iters 20 2>&1 | grep -v "^nParts\|k-way\|partition fin" | tee gpurun_out/rmat22.log
