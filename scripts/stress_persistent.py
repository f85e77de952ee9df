#!/usr/bin/env python
"""Development tool: randomized stress of the persistent kernel against the staged one
(bitwise when nothing overflows, else the accuracy gate), many (P, W, threads) combinations,
repeated back-to-back products."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ehyb_spmv_gpu_b200 import api
from oracle import oracle as O
from tests import util

orc = O.Oracle()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
cases = 0
t0 = time.time()
for kind, dims in (("st27", (40, 36, 33)), ("lap2d", (300, 211)), ("elas", (14, 12, 11))):
    n = util.lower_entries(kind, dims)[0]
    for _ in range(6):
        P = int(rng.integers(150, 1200))
        W = int(rng.integers(1, 8)) * 64 + (n // P // 64) * 64
        threads = int(rng.choice([0, 256, 512, 640, 768]))
        part = (np.arange(n, dtype=np.int64) * P // n).astype(np.uint32)
        rng.shuffle(part[: n // 50])            # ragged partitions (some tiny, some empty-ish)
        x = util.x_random(n, int(rng.integers(1 << 30)))
        m = util.product_pipeline(kind, dims, P, W, 1, x=x, partVec=np.sort(part) if rng.random() < 0.5 else part)
        lay = api.Layout(m, er_fill=float(rng.choice([-1.0, 0.0, 0.5])), cache_cap=int(rng.choice([0, 64, 16384])))
        st = lay.stats()
        s3 = api.Session(lay, kernel=api.KERNEL_PERSISTENT, threads=threads)
        s2 = api.Session(lay, kernel=api.KERNEL_STAGED, threads=threads)
        a = m.arrays()
        for rep in range(3):
            xr = m.vector_reorder(util.x_random(n, rep))
            y3, y2 = s3.spmv_host(xr), s2.spmv_host(xr)
            if st["nOverflow"] == 0:
                assert np.array_equal(y3, y2), (kind, P, W, threads, "bitwise")
            util.assert_within_gate(y3, orc.csr_spmv(a["rowIdx"], a["J"], a["V"], xr), orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], xr))
        s3.set_x(m.vector_reorder(x))
        s3.time_spmv(2, 30)
        util.assert_within_gate(s3.get_y(), orc.csr_spmv(a["rowIdx"], a["J"], a["V"], m.vector_reorder(x)), orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], m.vector_reorder(x)))
        print("ok", kind, "P", P, "W", W, "threads", threads, s3.kernel_name(), "ovf", st["nOverflow"], flush=True)
        cases += 1
        s2.free(); s3.free(); lay.free(); m.free()
print("stress ok: %d cases in %.0f s" % (cases, time.time() - t0))
