mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_driver.py tests/test_gpu_solver.py -m gpu -q -x --durations=8 ) > gpurun_out/r2_pytest_gpu_n2.log 2>&1; tail -30 gpurun_out/r2_pytest_gpu_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_bench_n2_512.json 2> gpurun_out/r2_bench_n2_512.err; tail -c 400 gpurun_out/r2_bench_n2_512.err; python -c "
import json
for l in open('gpurun_out/r2_bench_n2_512.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['config']['workload'][:70], d['parity'])"
