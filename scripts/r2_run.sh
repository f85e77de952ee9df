mkdir -p gpurun_out
( time timeout 1000 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -14 gpurun_out/r2_pytest_gpu_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 400 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n1_final.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ovfstream_kernel" --launch-skip 3 --launch-count 1 -o gpurun_out/r2_ovfstream_final_full -f python scripts/rmat_variants.py --scale 24 --iters 2 --variants default --no-cusparse > gpurun_out/ncu_ovf3.log 2>&1; tail -2 gpurun_out/ncu_ovf3.log
