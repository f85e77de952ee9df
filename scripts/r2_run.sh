mkdir -p gpurun_out
timeout 120 ./bin/spmv.out -g lap2d:1024:1024 -i 300 > gpurun_out/r2_spmv_out_c1.log 2>&1; tail -7 gpurun_out/r2_spmv_out_c1.log
timeout 120 ./bin/spmv.out -g st27:128:128:128 -i 200 > gpurun_out/r2_spmv_out_c2.log 2>&1; tail -5 gpurun_out/r2_spmv_out_c2.log
timeout 400 python bench.py > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n1_c.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']); print(d['comparisons'].get('config1_l2'))"
