mkdir -p gpurun_out
for B in 16x16x16 16x16x24 16x32x16 32x16x16 8x16x32; do
  EHYB_BENCH_GRID=256x256x256 EHYB_BENCH_BRICK=$B timeout 200 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_brick_$B.json 2> gpurun_out/r2_brick_$B.err
  python -c "
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/r2_brick_$B.json') if l.startswith('{')][-1]); print('$B', d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'], d['config'].get('partitions_rank0'), d['config'].get('window'), d['config'].get('remainder_cache_max'), d['parity'])
except Exception as e: print('$B failed', e); print(open('gpurun_out/r2_brick_$B.err').read()[-600:])
"
done
