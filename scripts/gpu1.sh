set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
python scripts/sweep.py --kind st27 --dims 128 128 128 --iters 200 --check --configs 296:0:1:512,296:0:1:1024,148:0:1:1024,148:0:2:512,592:0:1:256,444:0:1:320 2>&1 | grep -v "^nParts\|k-way\|partition fin\|partition time" | tee gpurun_out/sweep1.log
