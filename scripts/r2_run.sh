mkdir -p gpurun_out
T=tests/test_gpu_multi.py::test_distributed_product_on_gpus
( time timeout 900 python -m pytest "$T[4-persistent-p2p-grid-metis-48x40x36]" "$T[4-persistent-p2p-grid-scatter-40x40x24]" "$T[4-persistent-p2p-grid-metis-160x160x96]" "$T[4-staged-p2p-grid-scatter-40x40x24]" "$T[4-persistent-nccl-grid-metis-48x40x36]" "$T[4-persistent-p2p-oneway--]" "$T[4-staged-p2p-metis-64x64x24]" tests/test_gpu_driver.py::test_driver_multi_gpu_mode -m gpu -q -x --durations=4 ) > gpurun_out/r2_pytest_gpu_n4.log 2>&1; tail -14 gpurun_out/r2_pytest_gpu_n4.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/r2_bench_n4_512.json 2> gpurun_out/r2_bench_n4_512.err; tail -c 300 gpurun_out/r2_bench_n4_512.err; python -c "
import json
for l in open('gpurun_out/r2_bench_n4_512.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['config']['workload'][:70], d['parity'])"
