set -x
python - <<'PY'
import sys; sys.path.insert(0,'.')
from ehyb_spmv_gpu_b200 import api, _lib as L
import ctypes as C
n, li, lj, lv = api.gen_lower(api.GEN_LAPLACE2D, 1024, 1024)
lib = L.load()
lib.ehyb_write_mtx(b"read/lap2d_1024.mtx", n, C.c_int64(len(li)), li.ctypes.data_as(L.c_int_p), lj.ctypes.data_as(L.c_int_p), lv.ctypes.data_as(L.c_dbl_p), 1)
PY
for dd in 1 0; do for pdl in 1; do echo "DYNAMIC_DEAL=$dd"; EHYB_DYNAMIC_DEAL=$dd ./bin/spmv.out -i 2000 -m lap2d_1024 2>&1 | grep "EHYB-B200 events"; done; done
for u in 4 1; do for dd in 1 0; do echo "OVF_UNROLL=$u DYNAMIC_DEAL=$dd"; EHYB_OVF_UNROLL=$u EHYB_DYNAMIC_DEAL=$dd timeout 300 python scripts/run_rmat.py --scale 20 --blocks 2>&1 | grep "us per product\|main kernel"; done; done
echo "HINT=0"; EHYB_L2_HINT=0 timeout 300 python scripts/run_rmat.py --scale 20 --blocks 2>&1 | grep "us per product\|main kernel"
