set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys; sys.path.insert(0,'.')
from ehyb_spmv_gpu_b200 import api, _lib as L
import ctypes as C
n, li, lj, lv = api.gen_lower(api.GEN_LAPLACE2D, 1024, 1024)
lib = L.load()
lib.ehyb_write_mtx(b"read/lap2d_1024.mtx", n, C.c_int64(len(li)), li.ctypes.data_as(L.c_int_p), lj.ctypes.data_as(L.c_int_p), lv.ctypes.data_as(L.c_dbl_p), 1)
PY
./bin/spmv.out -i 2000 -m lap2d_1024 2>&1 | grep -v "^at \|large diff" > gpurun_out/spmv_out_c1.log; grep "EHYB-B200" gpurun_out/spmv_out_c1.log
timeout 300 python scripts/run_rmat.py --scale 20 2>&1 | grep -v "^start\|^partition fin\|nParts" | tee gpurun_out/rmat20.log
timeout 400 python scripts/run_rmat.py --scale 22 --blocks 2>&1 | tee gpurun_out/rmat22_blocks.log
timeout 200 python bench.py --steps 100 --warmup 5 2>/dev/null | tee gpurun_out/bench_n1_cmp.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], json.dumps(d['comparisons']))"
