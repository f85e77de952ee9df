set -x
python scripts/sweep.py --iters 20 --configs 148:0:1:0:0:2 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ehyb_staged -s 20 -c 1 -o gpurun_out/prof_r1_b python scripts/sweep.py --iters 20 --configs 148:0:1:0:0:2 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
