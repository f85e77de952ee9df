set -x
python - <<'PY'
import sys; sys.path.insert(0,'.')
from ehyb_spmv_gpu_b200 import api, _lib as L
import ctypes as C
n, li, lj, lv = api.gen_lower(api.GEN_LAPLACE2D, 1024, 1024)
lib = L.load()
lib.ehyb_write_mtx(b"read/lap2d_1024.mtx", n, C.c_int64(len(li)), li.ctypes.data_as(L.c_int_p), lj.ctypes.data_as(L.c_int_p), lv.ctypes.data_as(L.c_dbl_p), 1)
PY
./bin/spmv.out -i 2000 -m lap2d_1024 -C 2>&1 | grep -v "^at \|large diff" > gpurun_out/spmv_out_c1.log; grep "EHYB-B200: \|EHYB-B200 events\|rows fail" gpurun_out/spmv_out_c1.log
EHYB_KERNEL=2 ./bin/spmv.out -i 2000 -m lap2d_1024 -C -P 296 -W 3648 2>&1 | grep "EHYB-B200 events"
./bin/spmv.out -i 500 -g elas:100:100:100 2>&1 | grep -v "^at \|large diff" > gpurun_out/spmv_out_c3.log; grep "EHYB-B200: \|EHYB-B200 events\|rows fail" gpurun_out/spmv_out_c3.log
./bin/spmv.out -i 2000 -g st27:128:128:128 2>&1 | grep -v "^at \|large diff" > gpurun_out/spmv_out_c2.log; grep "EHYB-B200: \|EHYB-B200 events\|rows fail" gpurun_out/spmv_out_c2.log
