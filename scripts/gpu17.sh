set -x
python bench.py --steps 200 --warmup 10 2>/dev/null | tee gpurun_out/bench_n1_b.json | cut -c1-1500
# the reference's workflow: .mtx on disk, ./spmv.out -i 2000 -m <name>  (config 1)
python - <<'PY'
import sys; sys.path.insert(0,'.')
from ehyb_spmv_gpu_b200 import api, _lib as L
import ctypes as C
n, li, lj, lv = api.gen_lower(api.GEN_LAPLACE2D, 1024, 1024)
lib = L.load()
lib.ehyb_write_mtx(b"read/lap2d_1024.mtx", n, C.c_int64(len(li)), li.ctypes.data_as(L.c_int_p), lj.ctypes.data_as(L.c_int_p), lv.ctypes.data_as(L.c_dbl_p), 1)
PY
./bin/spmv.out -i 2000 -m lap2d_1024 2>&1 | tee gpurun_out/spmv_out_c1.log
./bin/spmv.out -i 2000 -g st27:128:128:128 2>&1 | tee gpurun_out/spmv_out_c2.log | tail -12
./bin/spmv.out -i 500 -g elas:100:100:100 2>&1 | tee gpurun_out/spmv_out_c3.log | tail -14
