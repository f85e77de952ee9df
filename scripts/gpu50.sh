set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k persistent 2>&1 | tail -3
EHYB_BENCH_GRID=256x256x256 EHYB_BENCH_SCALING=strong EHYB_KERNEL=3 EHYB_PARTS_PER_SM=28 timeout 900 python bench.py --steps 50 --warmup 5 2>gpurun_out/b256p_err.log | tee gpurun_out/bench_strong256_n1_persistent.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['config']['partitions'], d['config']['window'], d['roofline']['achieved'], d['parity'])"
grep "EHYB\|rror" gpurun_out/b256p_err.log | tail -3
