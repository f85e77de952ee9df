mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n1_final.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_isolated'], d['comparisons']['config1_l2']['us_per_product_l2_resident'])"
