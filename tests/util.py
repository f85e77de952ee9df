"""Shared helpers for the parity tests."""
from __future__ import annotations

import functools

import numpy as np

from ehyb_spmv_gpu_b200 import api
from oracle import oracle as O

KINDS = {"lap2d": (api.GEN_LAPLACE2D, O.gen_laplace2d_lower), "st27": (api.GEN_STENCIL27, O.gen_stencil27_lower),
         "elas": (api.GEN_ELASTICITY, O.gen_elasticity_lower)}

EHYB_KEYS = ["widthVecBlockELL", "biasVecBlockELL", "colBlockELL", "valBlockELL", "reorderListER", "rowVecER",
             "widthVecER", "biasVecER", "colER", "valER"]
COO_KEYS = ["I", "J", "V", "rowIdx", "numInRow", "numInRow2", "partBoundary", "reorderList"]


@functools.lru_cache(maxsize=None)
def lower_entries(kind: str, dims: tuple):
    """File entries of a symmetric generator matrix, from the numpy (test-side) generator."""
    return KINDS[kind][1](*dims)


@functools.lru_cache(maxsize=None)
def metis_partition(kind: str, dims: tuple, nparts: int):
    """Partition vector from the pinned mt-metis binary for the symmetric graph of the matrix."""
    orc = O.Oracle()
    n, li, lj, lv = lower_entries(kind, dims)
    m = orc.read_sym(n, li, lj, lv)
    xadj, adj = orc.graph(m)
    return O.mtmetis_partition(xadj, adj, nparts, nthreads=1)


def x_random(n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, n)


def product_pipeline(kind, dims, nParts, W, kpp=1, x=None, partVec=None):
    """generate -> matrixCOO -> reorder with the given (or mt-metis) partition, product code."""
    n, li, lj, lv = lower_entries(kind, dims)
    m = api.CooMatrix.from_lower(n, li, lj, lv, x)
    m.set_plan(nParts, W, kpp)
    m.reorder_with_partition(partVec if partVec is not None else metis_partition(kind, dims, nParts))
    return m


def oracle_pipeline(orc, kind, dims, nParts, W, x=None, partVec=None):
    n, li, lj, lv = lower_entries(kind, dims)
    mo = orc.read_sym(n, li, lj, lv, x)
    ro = orc.reorder(mo, nParts, W, partVec if partVec is not None else metis_partition(kind, dims, nParts))
    return mo, ro


def assert_within_gate(y, y_ref, absAx, tol=1e-12):
    """BASELINE.json accuracy gate: |y - y_ref| <= tol * (|A||x|) per row."""
    bad = np.flatnonzero(~(np.abs(y - y_ref) <= tol * absAx))
    assert bad.size == 0, f"{bad.size} rows outside the gate, first {bad[:5]}, err {np.abs(y - y_ref)[bad[:5]]}, bound {tol * absAx[bad[:5]]}"


def layout_spmv(raw, x_ext):
    """y = A x evaluated from the device-facing arrays of a tuned layout (Layout.raw()) with the
    index formulas of DESIGN.md section 3, written in numpy independently of layout.c and of the
    kernels.  x_ext has ncols entries ([own | halo] for a distributed block)."""
    parts, sl, blob = raw["parts"], raw["slices"], raw["blob"]
    n = int(parts[-1][1])
    y = np.zeros(n)
    for p in range(len(parts)):
        rs, re_, s0, s1, cs, cc = (int(v) for v in parts[p][:6])
        cache = raw["cacheCols"][cs:cs + cc]
        for s in range(s0, s1):
            off = int(sl["off256"][s]) * 256
            w, wr = int(sl["w"][s]), int(sl["wr"][s])
            w4, wr4 = (w + 3) // 4, (wr + 3) // 4
            ev = blob[off:off + w * 512].view(np.float64).reshape(w, 32, 2)
            ec = blob[off + w * 512:off + w * 512 + w4 * 512].view(np.uint16).reshape(w4, 32, 2, 4)
            ro = off + w * 512 + w4 * 512
            rv = blob[ro:ro + wr * 512].view(np.float64).reshape(wr, 32, 2)
            rc = blob[ro + wr * 512:ro + wr * 512 + wr4 * 512].view(np.uint16).reshape(wr4, 32, 2, 4)
            acc = np.zeros((32, 2))
            if w:
                cols = ec.transpose(0, 3, 1, 2).reshape(w4 * 4, 32, 2)[:w].astype(np.int64) + rs
                acc += (ev * x_ext[np.minimum(cols, len(x_ext) - 1)]).sum(axis=0)
            if wr:
                idx = rc.transpose(0, 3, 1, 2).reshape(wr4 * 4, 32, 2)[:wr].astype(np.int64)
                acc += (rv * x_ext[cache[idx]]).sum(axis=0)
            r0 = rs + (s - s0) * 64
            for h in range(2):
                lo = r0 + 32 * h
                cnt = max(0, min(32, re_ - lo))
                y[lo:lo + cnt] = acc[:cnt, h]
    np.add.at(y, raw["ovfRow"], raw["ovfVal"] * x_ext[raw["ovfCol"]])
    return y
