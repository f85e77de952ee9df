set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
python scripts/sweep.py --iters 200 --configs 148:0:2:512,148:0:2:512:0,148:0:2:512:1.0,148:0:4:256,148:0:4:512,148:0:3:320,148:0:2:384,148:0:2:256,222:0:2:512,296:0:2:512,296:0:2:256,296:0:4:256 2>&1 | grep "^cfg\|^#" | tee gpurun_out/sweep2.log
python scripts/sweep.py --iters 200 --no-l2 --configs 148:0:2:512,296:0:1:512 2>&1 | grep "^cfg\|^#" | tee gpurun_out/sweep2_nol2.log
python scripts/sweep.py --iters 20 --configs 148:0:2:512 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ehyb_main -s 20 -c 2 -o gpurun_out/prof_r1_a python scripts/sweep.py --iters 20 --configs 148:0:2:512 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
