"""Deterministic, parallel partition stage (csrc/host/hierpart.c) on CPU: a valid, balanced
partition vector, identical on every run, close to the single mt-metis call in quality, and the same
downstream parity (the product's reorder + format build from it equal the oracle's from the same vector)."""
import ctypes as C

import numpy as np
import pytest

from ehyb_spmv_gpu_b200 import _lib as L
from ehyb_spmv_gpu_b200 import api
from tests import util


def _hier(lib, xadj, adj, n, P, pieces):
    where = np.zeros(n, np.uint32)
    L.check(lib, lib.ehyb_partition_graph_hier(C.c_uint32(n), xadj.ctypes.data_as(L.c_u32_p), adj.ctypes.data_as(L.c_u32_p), C.c_uint32(P), pieces,
                                               where.ctypes.data_as(L.c_u32_p)), "ehyb_partition_graph_hier")
    return where


def test_hierarchical_partition_is_deterministic_and_balanced(lib, orc):
    kind, dims, P, W = "st27", (64, 64, 64), 120, 2304
    n, li, lj, lv = util.lower_entries(kind, dims)
    m = api.CooMatrix.from_lower(n, li, lj, lv)
    xadj, adj = m.build_graph()
    w1 = _hier(lib, xadj, adj, n, P, 4)
    w2 = _hier(lib, xadj, adj, n, P, 4)
    assert np.array_equal(w1, w2), "two runs gave different partitions"
    sizes = np.bincount(w1, minlength=P)
    # (a piece's share of the partitions is an integer: with 30 partitions per piece the sizes can be off by ~3 %)
    assert sizes.min() > 0 and sizes.max() <= 1.05 * n / P, (sizes.min(), sizes.max(), n / P)
    # quality: entries that leave their partition, against the reference's single call
    single = util.metis_partition(kind, dims, P)
    rows = np.repeat(np.arange(n), np.diff(xadj.astype(np.int64)))
    cut_h = int(np.count_nonzero(w1[rows] != w1[adj]))
    cut_s = int(np.count_nonzero(single[rows] != single[adj]))
    assert cut_h <= 1.10 * cut_s, (cut_h, cut_s)
    # downstream parity given this vector: product reorder + reference-layout convert == oracle's
    m.set_plan(P, W, 1)
    m.reorder_with_partition(w1)
    mo, ro = util.oracle_pipeline(orc, kind, dims, P, W, partVec=w1)
    a = m.arrays()
    for k in util.COO_KEYS:
        assert np.array_equal(a[k][:len(ro[k])] if k == "partBoundary" else a[k], ro[k][:len(a[k])] if k == "partBoundary" else ro[k]), k
    m.free()


def test_reorder_uses_the_stage_when_asked(lib):
    """ehyb_set_partition_pieces / $EHYB_PARTITION_PIECES switch ehyb_reorder (matrixReorder) over; small
    matrices (n < 65536) keep the reference's single call."""
    n, li, lj, lv = util.lower_entries("st27", (48, 48, 48))
    outs = []
    for pieces in (4, 4, 0):
        m = api.CooMatrix.from_lower(n, li, lj, lv)
        m.set_plan(30, 3840, 1)
        lib.ehyb_set_partition_pieces(pieces)
        m.reorder()
        outs.append(m.arrays()["reorderList"].copy())
        m.free()
    lib.ehyb_set_partition_pieces(0)
    assert np.array_equal(outs[0], outs[1])
    assert lib.ehyb_get_partition_pieces() == 0
