set -x
python bench.py --steps 200 --warmup 10 2>gpurun_out/b1_err.log | tee gpurun_out/bench_final_n1.json | cut -c1-300
python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tee gpurun_out/bench_final_ref.json | cut -c1-300
python bench.py --steps 20 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ehyb_staged -s 5 -c 2 -o gpurun_out/r1_staged_full python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python - <<'PY'
import sys; sys.path.insert(0,'.')
from ehyb_spmv_gpu_b200 import api, _lib as L
import ctypes as C
n, li, lj, lv = api.gen_lower(api.GEN_LAPLACE2D, 1024, 1024)
lib = L.load()
lib.ehyb_write_mtx(b"read/lap2d_1024.mtx", n, C.c_int64(len(li)), li.ctypes.data_as(L.c_int_p), lj.ctypes.data_as(L.c_int_p), lv.ctypes.data_as(L.c_dbl_p), 1)
PY
./bin/spmv.out -i 2000 -m lap2d_1024 2>&1 | grep -v "^at " > gpurun_out/spmv_out_c1.log; tail -6 gpurun_out/spmv_out_c1.log
./bin/spmv.out -i 2000 -g st27:128:128:128 2>&1 | grep -v "^at " > gpurun_out/spmv_out_c2.log; tail -5 gpurun_out/spmv_out_c2.log
./bin/spmv.out -i 500 -g elas:100:100:100 2>&1 | grep -v "^at " > gpurun_out/spmv_out_c3.log; tail -7 gpurun_out/spmv_out_c3.log
