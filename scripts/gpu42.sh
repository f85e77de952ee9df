python - <<'PY'
import sys; sys.path.insert(0,'.')
from ehyb_spmv_gpu_b200 import api, _lib as L
import ctypes as C
n, li, lj, lv = api.gen_lower(api.GEN_LAPLACE2D, 1024, 1024)
lib = L.load()
lib.ehyb_write_mtx(b"read/lap2d_1024.mtx", n, C.c_int64(len(li)), li.ctypes.data_as(L.c_int_p), lj.ctypes.data_as(L.c_int_p), lv.ctypes.data_as(L.c_dbl_p), 1)
PY
for cfg in "0 1" "1 1" "1 0" "0 0"; do set -- $cfg
  echo "== PROLOGUE_BARRIER=$1 DYNAMIC_DEAL=$2"
  EHYB_PROLOGUE_BARRIER=$1 EHYB_DYNAMIC_DEAL=$2 ./bin/spmv.out -i 2000 -m lap2d_1024 2>&1 | grep "EHYB-B200 events"
done
echo "== threads 512 / 384 (barrier 0, deal 0)"
EHYB_THREADS=512 EHYB_DYNAMIC_DEAL=0 ./bin/spmv.out -i 2000 -m lap2d_1024 2>&1 | grep "EHYB-B200 events"
EHYB_THREADS=384 EHYB_DYNAMIC_DEAL=0 ./bin/spmv.out -i 2000 -m lap2d_1024 2>&1 | grep "EHYB-B200 events"
