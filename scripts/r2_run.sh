mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "overflow or power_law or general" ) > gpurun_out/r2_pytest_ovf.log 2>&1; tail -15 gpurun_out/r2_pytest_ovf.log
timeout 400 python scripts/rmat_variants.py --scale 24 --iters 30 > gpurun_out/r2_rmat24_v2.log 2>&1; tail -14 gpurun_out/r2_rmat24_v2.log
timeout 400 python bench.py > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; tail -c 600 gpurun_out/r2_bench_n1_b.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n1_b.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['comparisons'].get('config1_l2'))"
