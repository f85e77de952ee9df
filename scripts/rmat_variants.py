#!/usr/bin/env python
"""BASELINE.json config 4 (R-MAT) timing of the overflow paths on one layout: the tile-packed CSR-like
stream (tile shapes, staging depth, hub columns), the COO list with atomics, cuSPARSE CSR on the same matrix.
  python scripts/rmat_variants.py --scale 24 [--iters 30] [--variants default,coo] [--no-cusparse]"""
import argparse, ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ehyb_spmv_gpu_b200 import api
from oracle import oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=22)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--variants", default="")
ap.add_argument("--no-cusparse", action="store_true")
ap.add_argument("--order", default="blocks", choices=["blocks", "degree"],
                help="level-1 numbering: contiguous blocks of the generator's numbering, or columns by descending reference count (the hubs first)")
a = ap.parse_args()
t = time.time()
n, fi, fj, fv = api.gen_rmat(a.scale, 16, seed=1, add_diagonal=False)
x = np.random.default_rng(0).uniform(-0.1, 0.1, n)
m = api.CooMatrix.from_general(n, fi, fj, fv, x)
if a.order == "degree":
    # 4 096-vertex partitions in descending order of the column reference counts: the permuted numbering
    # starts with the hubs, so the x entries most gathers go to share a few hundred cache lines
    cnt = np.bincount(fj, minlength=n)
    rank = np.empty(n, np.int64)
    rank[np.argsort(-cnt, kind="stable")] = np.arange(n)
    P = max(1, n // 4096)
    m.set_plan(P, 4096, 1)
    part = (rank * P // n).astype(np.uint32)
    del cnt, rank
else:
    pl = api.plan(n)
    m.set_plan(pl.nParts, pl.W, pl.ctasPerPart)
    part = (np.arange(n, dtype=np.int64) * pl.nParts // n).astype(np.uint32)
del fi, fj, fv
m.reorder_with_partition(part)
lay = api.Layout(m)
st = lay.stats()
print(f"rmat scale {a.scale}, order {a.order}: n={n} nnz={st['nnz']} host {time.time()-t:.1f}s; overflow {st['nnzOverflow']} algBytes {st['algBytes']} COO formatBytes {st['formatBytes']}", flush=True)
xr = m.vector_reorder(x)
orc = O.Oracle()
arr = m.arrays()
y_ref = orc.csr_spmv(arr["rowIdx"], arr["J"], arr["V"], xr)
absAx = orc.csr_abs_spmv(arr["rowIdx"], arr["J"], arr["V"], xr)
VARIANTS = (("default", "stream: 128-entry tiles x 32 warps, 2 slots, 2048 hubs, 1 column block", {}),
            ("nohubs", "stream: no hubs", {"EHYB_OVF_HUBS": "0"}),
            ("k2", "stream: 2 column blocks", {"EHYB_OVF_COLBLOCKS": "2"}),
            ("k4", "stream: 4 column blocks", {"EHYB_OVF_COLBLOCKS": "4"}),
            ("k8", "stream: 8 column blocks", {"EHYB_OVF_COLBLOCKS": "8"}),
            ("hubs1k", "stream: 1024 hubs", {"EHYB_OVF_HUBS": "1024"}),
            ("hubs2k", "stream: 2048 hubs", {"EHYB_OVF_HUBS": "2048"}),
            ("hubs4k", "stream: 4096 hubs", {"EHYB_OVF_HUBS": "4096"}),
            ("hubs6k", "stream: 6144 hubs", {"EHYB_OVF_HUBS": "6144"}),
            ("hubs12k", "stream: 12288 hubs", {"EHYB_OVF_HUBS": "12288"}),
            ("hubs8k", "stream: 8192 hubs", {"EHYB_OVF_HUBS": "8192"}),
            ("slots3", "stream: 3 slots", {"EHYB_OVF_SLOTS": "3"}),
            ("slots4", "stream: 4 slots", {"EHYB_OVF_SLOTS": "4"}),
            ("coo", "COO list + atomics (round 1)", {"EHYB_OVF_STREAM": "0"}))
want = [w for w in a.variants.split(",") if w]
for key, name, env in VARIANTS:
    if want and key not in want:
        continue
    for k in ("EHYB_OVF_HUBS", "EHYB_OVF_STREAM", "EHYB_OVF_TG", "EHYB_OVF_SLOTS", "EHYB_OVF_COLBLOCKS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    s = api.Session(lay)
    y = s.spmv_host(xr)
    bad = int(np.count_nonzero(~(np.abs(y - y_ref) <= 1e-12 * absAx)))
    same = bool(np.array_equal(s.spmv_host(xr), y))
    s.set_x(xr)
    ms = s.time_spmv(5, a.iters)
    ms = ms[0] if isinstance(ms, tuple) else ms
    per = ms / a.iters
    print(f"{name:52s} {per*1e3:8.1f} us  {2*st['nnz']/(per*1e6):7.1f} GFLOP/s  {st['algBytes']/(per*1e6):7.1f} GB/s alg  gate fails {bad}  bit-reproducible {same}  kernel {s.kernel_name()} launches {s.launches_per_spmv()}", flush=True)
    s.free()
cus = os.path.join(ROOT, "ehyb_spmv_gpu_b200", "lib", "libehyb_cusparse.so")
if os.path.exists(cus) and not a.no_cusparse:
    lib = C.CDLL(cus)
    for alg in (1, 2):
        us = C.c_float(); yc = np.empty(n)
        if lib.ehyb_cusparse_spmv(C.byref(m.c), xr.ctypes.data_as(C.POINTER(C.c_double)), yc.ctypes.data_as(C.POINTER(C.c_double)), 3, 20, alg, C.byref(us)) == 0:
            print(f"cuSPARSE CSR ALG{alg}: {us.value:.1f} us per product, {2*st['nnz']/(us.value*1e3):.1f} GFLOP/s", flush=True)
