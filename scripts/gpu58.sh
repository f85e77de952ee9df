set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ehyb_persistent -s 5 -c 2 -o gpurun_out/r1_persistent_full python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python bench.py --steps 200 --warmup 10 2>/dev/null | tee gpurun_out/bench_final_n1.json | cut -c1-200
