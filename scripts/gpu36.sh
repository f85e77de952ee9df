set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/run_rmat.py --scale 20 --blocks --iters 20 2>&1 | tee gpurun_out/rmat20_blocks.log | grep "layout\|us per product\|main kernel\|gate"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/rmat20_launches.csv python scripts/run_rmat.py --scale 20 --blocks --iters 5 > gpurun_out/ncu_rmat.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ehyb_overflow -s 3 -c 1 -o gpurun_out/r1_overflow_full python scripts/run_rmat.py --scale 20 --blocks --iters 5 > gpurun_out/ncu_rmat_full.log 2>&1
tail -2 gpurun_out/ncu_rmat_full.log
