mkdir -p gpurun_out
P='import json,sys; t=sys.stdin.read(); d=json.loads(t[t.index("{\"metric"):]); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], list(d["parity"].values())[:2])'
A=$PWD/ehyb_spmv_gpu_b200/lib/libehyb_A.so
run() { echo "== 256^3 $*"; env "$@" EHYB_PERSIST_SLOTS=2 EHYB_BENCH_GRID=256x256x256 timeout 600 python bench.py --steps 200 --warmup 5 2>gpurun_out/err_b.log | python -c "$P" || tail -3 gpurun_out/err_b.log; }
runc2() { echo "== C2 $*"; env "$@" timeout 600 python bench.py --steps 200 --warmup 10 2>gpurun_out/err_b.log | python -c "$P" || tail -3 gpurun_out/err_b.log; }
runc2 EHYB_LIB=$A
runc2 B=1
runc2 EHYB_LIB=$A EHYB_THREADS=512
runc2 EHYB_THREADS=512
run EHYB_LIB=$A EHYB_THREADS=512
run EHYB_THREADS=512
run EHYB_LIB=$A EHYB_THREADS=640
run EHYB_THREADS=640
