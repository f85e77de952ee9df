#!/usr/bin/env python
"""BASELINE.json config 4: R-MAT power-law graph through the general (unsymmetric) path.
  python scripts/run_rmat.py --scale 22 [--edge-factor 16] [--iters 50]
The reference cannot run this input at any scale: the oracle is the CPU
CSR product."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("EHYB_MTMETIS_BIN", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bin", "ehyb_mtmetis"))
from ehyb_spmv_gpu_b200 import api
from oracle import oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=20)
ap.add_argument("--edge-factor", type=int, default=16)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--blocks", action="store_true", help="contiguous-block partition instead of mt-metis")
ap.add_argument("--no-graph", action="store_true")
a = ap.parse_args()
t = time.time()
n, fi, fj, fv = api.gen_rmat(a.scale, a.edge_factor, seed=1, add_diagonal=False)
print(f"rmat scale {a.scale}: n={n} nnz={len(fi)} gen {time.time()-t:.1f}s", flush=True)
x = api.x_reference(n) if n <= (1 << 22) else np.random.default_rng(0).uniform(-0.1, 0.1, n)
m = api.CooMatrix.from_general(n, fi, fj, fv, x)
pl = api.plan(n)
m.set_plan(pl.nParts, pl.W, pl.ctasPerPart)
t = time.time()
if a.blocks:
    m.reorder_with_partition((np.arange(n, dtype=np.int64) * pl.nParts // n).astype(np.uint32))
else:
    m.reorder()
print(f"reorder {time.time()-t:.1f}s  P={pl.nParts} W={pl.W}", flush=True)
t = time.time()
lay = api.Layout(m)
st = lay.stats()
print(f"layout {time.time()-t:.1f}s", {k: st[k] for k in ("nSlices", "nnzEll", "nnzRemInSlice", "nnzOverflow", "padEll", "padRem", "nLongRows", "cacheMax", "algBytes", "formatBytes")}, flush=True)
s = api.Session(lay, use_graph=not a.no_graph)
xr = m.vector_reorder(x)
y = m.vector_recover(s.spmv_host(xr))
orc = O.Oracle()
arr = m.arrays()
y_ref = m.vector_recover(orc.csr_spmv(arr["rowIdx"], arr["J"], arr["V"], xr))
absAx = m.vector_recover(orc.csr_abs_spmv(arr["rowIdx"], arr["J"], arr["V"], xr))
bad = int(np.count_nonzero(~(np.abs(y - y_ref) <= 1e-12 * absAx)))
print("rows outside the 1e-12 gate:", bad, " max abs err", float(np.abs(y - y_ref).max()), " vs golden", float(np.abs(y - m.y_golden).max()))
s.set_x(xr)
ms, kms = s.time_spmv(5, a.iters, kernel_only=True)
per = ms / a.iters
print(f"main kernel alone {kms/a.iters*1e3:.1f} us (isolated launches)")
print(f"{per*1e3:.1f} us per product, {2*st['nnz']/(per*1e6):.1f} GFLOP/s, {st['algBytes']/(per*1e6):.1f} GB/s algorithmic, launches/product {s.launches_per_spmv()}")
# cuSPARSE CSR on the same permuted matrix, same GPU (comparison row)
import ctypes as C
cus = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ehyb_spmv_gpu_b200", "lib", "libehyb_cusparse.so")
if os.path.exists(cus):
    lib = C.CDLL(cus)
    for alg in (1, 2):
        us = C.c_float(); yc = np.empty(n)
        rc = lib.ehyb_cusparse_spmv(C.byref(m.c), xr.ctypes.data_as(C.POINTER(C.c_double)), yc.ctypes.data_as(C.POINTER(C.c_double)), 3, 20, alg, C.byref(us))
        if rc == 0:
            print(f"cuSPARSE CSR ALG{alg}: {us.value:.1f} us per product, {2*st['nnz']/(us.value*1e3):.1f} GFLOP/s")
sec, _ = orc.csr_spmv_timed(arr["rowIdx"], arr["J"], arr["V"], xr, 1, 5)
print(f"CPU CSR ({orc.num_threads()} threads): {2*st['nnz']*5/sec/1e9:.2f} GFLOP/s")
