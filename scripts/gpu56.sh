set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py 2>/dev/null | tee gpurun_out/bench_head_n1.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('C2 bench:', d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['config']['partitions'], d['config']['window'], d['parity'])"
./bin/spmv.out -i 2000 -m lap2d_1024 -C 2>&1 | grep "EHYB-B200: \|EHYB-B200 events\|rows fail"
./bin/spmv.out -i 500 -g elas:100:100:100 2>&1 | grep "EHYB-B200: \|EHYB-B200 events\|rows fail"
EHYB_BENCH_GRID=256x256x256 EHYB_BENCH_SCALING=strong timeout 900 python bench.py --steps 50 --warmup 5 2>/dev/null | tee gpurun_out/bench_strong256_n1_auto.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('256^3:', d['ms_per_step'], d['value'], d['roofline']['kernel'], d['config']['partitions'], d['config']['window'], d['config']['nnz_overflow'], d['parity'])"
