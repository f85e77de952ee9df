mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r2_bench_n8_512.json 2> gpurun_out/r2_bench_n8_512.err; tail -c 300 gpurun_out/r2_bench_n8_512.err; python -c "
import json
for l in open('gpurun_out/r2_bench_n8_512.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['config']['peers_per_gpu_min'], d['config']['peers_per_gpu_max'], d['config']['halo_x_entries_max_per_gpu'], d['parity'])"
timeout 200 ./bin/spmv.out -G 8 -g st27:512:512:512 -i 100 > gpurun_out/r2_spmv_out_G8_512.log 2>&1; tail -12 gpurun_out/r2_spmv_out_G8_512.log | cut -c1-250
