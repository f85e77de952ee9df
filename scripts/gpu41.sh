python - <<'PY'
import sys; sys.path.insert(0,'.')
from ehyb_spmv_gpu_b200 import api, _lib as L
import ctypes as C
n, li, lj, lv = api.gen_lower(api.GEN_LAPLACE2D, 1024, 1024)
lib = L.load()
lib.ehyb_write_mtx(b"read/lap2d_1024.mtx", n, C.c_int64(len(li)), li.ctypes.data_as(L.c_int_p), lj.ctypes.data_as(L.c_int_p), lv.ctypes.data_as(L.c_dbl_p), 1)
PY
for sha in d818cdb 19f2478 c95113a HEAD; do
  echo "== $sha"
  if [ $sha = HEAD ]; then ./bin/spmv.out -i 2000 -m lap2d_1024 2>&1 | grep "EHYB-B200 events"; else LD_LIBRARY_PATH=$PWD/build/bisect/$sha ./bin/spmv.out -i 2000 -m lap2d_1024 2>&1 | grep "EHYB-B200 events"; fi
done
