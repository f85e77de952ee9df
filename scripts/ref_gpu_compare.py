#!/usr/bin/env python
"""Reference GPU kernels (kernel.cu, unmodified, compiled for sm_100a by oracle/Makefile refgpu)
next to this engine on the same B200, same matrix, same x.  Measurement script, not a bench.

  python scripts/ref_gpu_compare.py [--kind st27 --dims 128 128 128] [--iters 200]
"""
import argparse, json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("EHYB_MTMETIS_BIN", os.path.join(ROOT, "bin", "ehyb_mtmetis"))
from oracle import oracle as O
from ehyb_spmv_gpu_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="st27")
ap.add_argument("--dims", type=int, nargs="+", default=[128, 128, 128])
ap.add_argument("--iters", type=int, default=200)
a = ap.parse_args()
gen = {"lap2d": O.gen_laplace2d_lower, "st27": O.gen_stencil27_lower, "elas": O.gen_elasticity_lower}[a.kind]
orc = O.Oracle()
n, li, lj, lv = gen(*a.dims)
x = orc.x_reference(n)
m = orc.read_sym(n, li, lj, lv, x)
out = {"matrix": f"{a.kind} {a.dims}", "n": n, "nnz": int(m["nnz"])}

# ---- reference: its own partition parameters (82-SM heuristic), its own COO2EHYB and kernels
P, W, kpp = orc.heuristic_ref(n, True)
xadj, adj = orc.graph(m)
part = O.mtmetis_partition(xadj, adj, P, nthreads=1)
r = orc.reorder(m, P, W, part)           # bit-identical to reference matrixReorder (tests/test_oracle_pinned.py)
xr = orc.vector_reorder(x, r["reorderList"])
with tempfile.TemporaryDirectory() as d:
    fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
    with open(fin, "wb") as f:
        f.write(np.array([n, m["nnz"], P, W, max(kpp, 1)], np.int32).tobytes())
        for k in ("I", "J"): f.write(r[k].tobytes())
        f.write(r["V"].tobytes())
        for k in ("rowIdx", "numInRow", "numInRow2", "partBoundary"): f.write(r[k].astype(np.int32).tobytes())
        f.write(xr.tobytes())
    res = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_gpu_bench"), fin, fout, str(a.iters)], capture_output=True, text=True)
    line = [l for l in res.stdout.splitlines() if l.startswith("{")]
    if res.returncode or not line:
        print(res.stdout[-2000:], res.stderr[-2000:]); sys.exit(1)
    out.update(json.loads(line[-1]))
    yy = np.fromfile(fout, dtype=np.float64)
y_ref = orc.csr_spmv(r["rowIdx"], r["J"], r["V"], xr)
absAx = orc.csr_abs_spmv(r["rowIdx"], r["J"], r["V"], xr)
for name, y in (("as_shipped", yy[:n]), ("repaired", yy[n:])):
    out["reference_gpu"]["rows_outside_gate_" + name] = int(np.count_nonzero(~(np.abs(y - y_ref) <= 1e-12 * absAx)))
e = orc.convert(r)
y_emul = orc.emulate(e, r, xr, use_fma=True)
out["reference_gpu"]["bit_identical_to_oracle_fma_emulation"] = bool(np.array_equal(yy[n:], y_emul))

# ---- this engine: B200 plan
pl = api.plan(n, api.device_query(0))
mm = api.CooMatrix.from_lower(n, li, lj, lv, x)
mm.set_plan(pl.nParts, pl.W, pl.ctasPerPart)
mm.reorder()
lay = api.Layout(mm)
s = api.Session(lay)
xr2 = mm.vector_reorder(x)
s.set_x(xr2)
ms = s.time_spmv(10, a.iters)
y2 = mm.vector_recover(s.get_y())
st = lay.stats()
out["this_engine"] = {"nParts": pl.nParts, "W": pl.W, "us_per_product": round(ms / a.iters * 1e3, 3),
                      "gflops": round(2.0 * st["nnz"] * a.iters / (ms * 1e6), 2),
                      "max_abs_err_vs_golden": float(np.abs(y2 - m["y"]).max())}
out["speedup_vs_reference_repaired"] = round(out["reference_gpu"]["us_per_product_repaired"] / out["this_engine"]["us_per_product"], 3)
out["speedup_vs_reference_as_shipped"] = round(out["reference_gpu"]["us_per_product_as_shipped"] / out["this_engine"]["us_per_product"], 3)
print(json.dumps(out))
