mkdir -p gpurun_out
P='import json,sys; t=sys.stdin.read(); d=json.loads(t[t.index("{\"metric"):]); print(d["ms_per_step"], d["value"], d["roofline"]["frac"])'
echo "== C2 B"; timeout 300 python bench.py --steps 200 --warmup 10 2>/dev/null | python -c "$P"
echo "== 200^3 B"; EHYB_BENCH_GRID=192x192x192 timeout 600 python bench.py --steps 100 --warmup 5 2>/dev/null | python -c "$P"
echo "== 256^3 B"; EHYB_BENCH_GRID=256x256x256 timeout 600 python bench.py --steps 200 --warmup 5 2>/dev/null | python -c "$P"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:persistent -s 5 -c 1 -f -o gpurun_out/r2_c2_persistent_full python bench.py --steps 10 --warmup 3 > gpurun_out/ncu2.log 2>&1
EHYB_BENCH_GRID=256x256x256 ncu --set full --clock-control none --import-source on -k regex:persistent -s 3 -c 1 -f -o gpurun_out/r2_256_persistent_full python bench.py --steps 5 --warmup 3 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
