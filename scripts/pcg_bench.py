#!/usr/bin/env python
"""Development/measurement tool: ehyb_pcg_solve on BASELINE.json config 2 (27-point 128^3).
Prints iterations, time per iteration and the share of the product in it."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("EHYB_MTMETIS_BIN", os.path.join(ROOT, "bin", "ehyb_mtmetis"))
from ehyb_spmv_gpu_b200 import api
from bench import build_matrix, stdout_to_stderr

with stdout_to_stderr():
    m, lay, x, pl, _ = build_matrix((128, 128, 128))
s = api.Session(lay)
a = m.arrays()
n = m.n
xr = m.vector_reorder(x)
b = s.spmv_host(xr)                     # b = A x_true
rows = np.repeat(np.arange(n), np.diff(a["rowIdx"]))
on = a["J"] == rows
diag = np.zeros(n); diag[rows[on]] = a["V"][on]
s.set_x(xr)
ms_spmv = s.time_spmv(10, 200); ms_spmv = (ms_spmv[0] if isinstance(ms_spmv, tuple) else ms_spmv) / 200
for jacobi in (True, False):
    for rtol in (1e-8,):
        xs, info = s.pcg_solve(b, diag if jacobi else None, max_iters=5000, rtol=rtol)
        err = np.linalg.norm(xs - xr) / np.linalg.norm(xr)
        print("PCG config 2 %s rtol %.0e: %d iterations, %.2f ms, %.1f us/iteration (product alone %.1f us = %.0f %%), true residual %.2e, |x - x_true|/|x_true| %.2e"
              % ("Jacobi" if jacobi else "plain ", rtol, info["iters"], info["ms"], info["ms"] * 1e3 / max(info["iters"], 1), ms_spmv * 1e3,
                 100 * ms_spmv * 1e3 / (info["ms"] * 1e3 / max(info["iters"], 1)), info["true_rel_residual"], err), flush=True)
