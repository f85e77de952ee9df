set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py 2>/dev/null | tee gpurun_out/bench_head_n1.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('staged default:', d['ms_per_step'], d['value'], d['roofline']['frac'], d['parity'])"
EHYB_KERNEL=3 EHYB_PARTS_PER_SM=3 timeout 300 python bench.py 2>/dev/null | tee gpurun_out/bench_head_n1_persistent.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('persistent P=444:', d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['partitions'], d['config']['nnz_overflow'], d['parity'])"
python - <<'PY'
import sys; sys.path.insert(0,'.')
from ehyb_spmv_gpu_b200 import api, _lib as L
import ctypes as C
n, li, lj, lv = api.gen_lower(api.GEN_LAPLACE2D, 1024, 1024)
lib = L.load()
lib.ehyb_write_mtx(b"read/lap2d_1024.mtx", n, C.c_int64(len(li)), li.ctypes.data_as(L.c_int_p), lj.ctypes.data_as(L.c_int_p), lv.ctypes.data_as(L.c_dbl_p), 1)
PY
./bin/spmv.out -i 2000 -m lap2d_1024 -C 2>&1 | grep "EHYB-B200 events\|rows fail"
