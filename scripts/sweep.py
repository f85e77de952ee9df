#!/usr/bin/env python
"""Parameter sweep on one GPU: partitions x window x CTAs/partition x threads for a stencil
matrix.  Development tool (not part of the product, not a bench): prints one line per config.

  python scripts/sweep.py --kind st27 --dims 128 128 128 --configs 296:0:1:512,148:0:1:1024
config = nParts:W:ctasPerPart:threads[:er_fill[:kernel]]  kernel 1=direct 2=staged   (W=0: smallest multiple of 64 >= largest partition)
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ehyb_spmv_gpu_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="st27")
    ap.add_argument("--dims", type=int, nargs="+", default=[128, 128, 128])
    ap.add_argument("--configs", default="296:0:1:512")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--no-l2", action="store_true")
    a = ap.parse_args()
    kind = {"lap2d": api.GEN_LAPLACE2D, "st27": api.GEN_STENCIL27, "elas": api.GEN_ELASTICITY}[a.kind]
    dims = a.dims + [1] * (3 - len(a.dims))
    t = time.time()
    n, li, lj, lv = api.gen_lower(kind, *dims)
    x = api.x_reference(n)
    print(f"# generated n={n} lower={len(li)} in {time.time()-t:.1f}s", flush=True)
    parts_cache = {}
    for cfg in a.configs.split(","):
        f = cfg.split(":")
        P, W, kpp, threads = int(f[0]), int(f[1]), int(f[2]), int(f[3])
        fill = float(f[4]) if len(f) > 4 and f[4] != "" else 0.5
        kern = int(f[5]) if len(f) > 5 and f[5] != "" else 0
        if len(f) > 6:
            os.environ["EHYB_CHUNK"] = f[6]
        t = time.time()
        m = api.CooMatrix.from_lower(n, li, lj, lv, x)
        m.set_plan(P, W if W else 64, kpp)
        if P in parts_cache:
            m.reorder_with_partition(parts_cache[P])
        else:
            xa, ad = m.build_graph()
            pv = np.zeros(n, np.uint32)
            from ehyb_spmv_gpu_b200 import _lib as L
            import ctypes as C
            lib = L.load()
            L.check(lib, lib.ehyb_partition_graph(C.c_uint32(n), xa.ctypes.data_as(L.c_u32_p), ad.ctypes.data_as(L.c_u32_p),
                                                  C.c_uint32(P), C.c_uint32(int(os.environ.get("SWEEP_METIS_THREADS", "1"))),
                                                  pv.ctypes.data_as(L.c_u32_p)), "partition")
            parts_cache[P] = pv
            m.reorder_with_partition(pv)
        arr = m.arrays()
        maxpart = int(np.diff(arr["partBoundary"]).max())
        Wuse = W if W else (maxpart + 63) // 64 * 64
        lay = api.Layout(m, W=Wuse, ctasPerPart=kpp, er_fill=fill)
        st = lay.stats()
        tprep = time.time() - t
        s = api.Session(lay, threads=threads, l2_persist_x=not a.no_l2, kernel=kern)
        xr = m.vector_reorder(x)
        s.set_x(xr)
        ms, kms = s.time_spmv(10, a.iters, kernel_only=True)
        per = ms / a.iters
        gbs = st["algBytes"] / (per * 1e6)
        kgbs = st["algBytes"] / (kms / a.iters * 1e6)
        ok = ""
        if a.check:
            y = m.vector_recover(s.get_y())
            err = np.abs(y - m.y_golden).max()
            ok = f" maxerr={err:.2e}"
        print(f"cfg {cfg:22s} W={Wuse:6d} maxpart={maxpart:6d} ell={st['nnzEll']/st['nnz']:.4f} ovf={st['nnzOverflow']} "
              f"fmt/alg={st['formatBytes']/st['algBytes']:.4f} us={per*1e3:8.2f} GB/s={gbs:7.1f} kernel_us={kms/a.iters*1e3:8.2f} "
              f"kGB/s={kgbs:7.1f} GF={2*st['nnz']/(per*1e6):7.1f} prep={tprep:.1f}s{ok}", flush=True)
        s.free(); lay.free(); m.free()


if __name__ == "__main__":
    main()
