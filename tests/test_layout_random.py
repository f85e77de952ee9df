"""Randomized host parity of the tuned layout (csrc/host/layout.c): random symmetric sparse
matrices with a few dense rows, random partitions / windows / options.  For every case:
  - the device-facing arrays, decoded with the formulas of DESIGN.md section 3 in numpy
    (tests/util.py: layout_spmv), give A x inside the accuracy gate;
  - the entry counts of the three classes add up;
  - the de-interleave reproduces the reference's EHYB arrays byte for byte - against the product's
    own COO2EHYB (itself pinned to the unmodified reference in test_host_parity.py) and, where the
    reference's converter accepts the input, against the oracle's restatement;
  - save -> load leaves all of it unchanged."""
import numpy as np
import pytest

from ehyb_spmv_gpu_b200 import api
from tests import util


def _random_lower(rng, n, avg, dense_rows):
    """lower-triangle entries (file order: by column, then row) of a random symmetric matrix"""
    cnt = int(n * avg / 2)
    i = rng.integers(0, n, cnt); j = rng.integers(0, n, cnt)
    lo, hi = np.minimum(i, j), np.maximum(i, j)
    keep = lo != hi
    lo, hi = lo[keep], hi[keep]
    for r in dense_rows:                                  # a few rows/columns touching many others
        others = rng.choice(n, size=min(n - 1, int(rng.integers(n // 3, n))), replace=False)
        others = others[others != r]
        lo = np.concatenate([lo, np.minimum(others, r)]); hi = np.concatenate([hi, np.maximum(others, r)])
    key = np.unique(lo.astype(np.int64) * n + hi)         # column-major: column = lo, row = hi
    col, row = (key // n).astype(np.int32), (key % n).astype(np.int32)
    d = np.arange(n, dtype=np.int32)
    li = np.concatenate([row, d]); lj = np.concatenate([col, d])
    order = np.lexsort((li, lj))
    li, lj = li[order], lj[order]
    lv = np.where(li == lj, 10.0 + rng.random(len(li)), rng.uniform(-1, 1, len(li)))
    return li.astype(np.int32), lj.astype(np.int32), lv


@pytest.mark.parametrize("seed", range(12))
def test_random_matrices_and_options(orc, seed, tmp_path):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(300, 2500))
    li, lj, lv = _random_lower(rng, n, avg=float(rng.uniform(3, 14)), dense_rows=rng.choice(n, size=int(rng.integers(0, 3)), replace=False))
    x = util.x_random(n, seed)
    m = api.CooMatrix.from_lower(n, li, lj, lv, x)
    P = int(rng.integers(1, 9))
    W = 32 * int(rng.integers(max(1, n // P // 64), max(2, n // P // 16) + 1))
    part = np.sort(rng.integers(0, P, n)).astype(np.uint32)
    if rng.random() < 0.5:
        part = rng.permutation(part)                      # partitions that are not index ranges
    m.set_plan(P, W, 1)
    m.reorder_with_partition(part)
    a = m.arrays()
    xr = m.vector_reorder(x)
    y_ref = orc.csr_spmv(a["rowIdx"], a["J"], a["V"], xr)
    absAx = orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], xr)
    e = m.coo2ehyb()                                      # the reference layout, product converter
    for trial in range(4):
        opts = dict(er_fill=float(rng.choice([-1.0, 0.0, 0.5, 1.0])), cache_cap=int(rng.choice([0, -1, 16, 200, 16384])),
                    long_row_threshold=int(rng.choice([0, 0, 40, 512])), min_coverage=float(rng.choice([0.0, -1.0, 0.6])))
        lay = api.Layout(m, **opts)
        st, raw = lay.stats(), lay.raw()
        assert st["nnzEll"] + st["nnzRemInSlice"] + st["nnzOverflow"] == st["nnz"] == m.nnz, opts
        assert st["nOverflow"] == st["nnzOverflow"] and np.all(np.diff(raw["ovfRow"]) >= 0)
        util.assert_within_gate(util.layout_spmv(raw, xr), y_ref, absAx)
        e2 = lay.to_reference()
        for k in util.EHYB_KEYS:
            assert np.array_equal(e2[k], e[k]), (opts, k)
        if trial == 0:
            path = tmp_path / "l.ehyb"
            lay.save(path)
            back = api.Layout.load(path)
            assert back.stats() == st
            rb = back.raw()
            for k in raw:
                assert np.array_equal(raw[k], rb[k]), k
            back.free()
        lay.free()
    m.free()


@pytest.mark.parametrize("seed", range(6))
def test_random_unsymmetric_matrices(orc, seed):
    """The general (unsymmetric) reader path with the same checks (reference solver_test.c:31-126)."""
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.integers(200, 1500))
    cnt = int(n * rng.uniform(2, 10))
    key = np.unique(rng.integers(0, n, cnt).astype(np.int64) * n + rng.integers(0, n, cnt))
    key = np.union1d(key, np.arange(n, dtype=np.int64) * (n + 1))          # full diagonal
    fi, fj = (key // n).astype(np.int32), (key % n).astype(np.int32)
    perm = rng.permutation(len(fi))                                        # file order is arbitrary
    fi, fj = fi[perm], fj[perm]
    fv = rng.uniform(-1, 1, len(fi))
    x = util.x_random(n, seed)
    m = api.CooMatrix.from_general(n, fi, fj, fv, x)
    P = int(rng.integers(1, 7))
    W = 32 * int(rng.integers(max(1, n // P // 64), max(2, n // P // 16) + 1))
    m.set_plan(P, W, 1)
    m.reorder_with_partition(rng.integers(0, P, n).astype(np.uint32))
    a = m.arrays()
    xr = m.vector_reorder(x)
    y_ref = orc.csr_spmv(a["rowIdx"], a["J"], a["V"], xr)
    absAx = orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], xr)
    # the reader's golden product (file-order accumulation) agrees with the permuted CSR product
    util.assert_within_gate(m.vector_recover(y_ref), m.y_golden, m.vector_recover(absAx))
    e = m.coo2ehyb()
    for opts in (dict(), dict(er_fill=0.0, cache_cap=16384), dict(cache_cap=-1), dict(min_coverage=0.9), dict(long_row_threshold=8)):
        lay = api.Layout(m, **opts)
        st = lay.stats()
        assert st["nnzEll"] + st["nnzRemInSlice"] + st["nnzOverflow"] == st["nnz"] == m.nnz
        util.assert_within_gate(util.layout_spmv(lay.raw(), xr), y_ref, absAx)
        e2 = lay.to_reference()
        for k in util.EHYB_KEYS:
            assert np.array_equal(e2[k], e[k]), (opts, k)
        lay.free()
    m.free()
