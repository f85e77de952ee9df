"""Streamed format build (BASELINE.json config 5's host pipeline) on CPU.

  * ehyb_layout_builder_* fed a few partitions at a time gives the layout of the one-shot build,
    byte for byte;
  * a brick-decomposed stencil block streamed by ehyb_mg_grid_build equals what the general path
    (ehyb_mg_local_build + ehyb_mg_local_finish on all rows, partition vector = brick) builds:
    permutation, halo list, every layout array;
  * the product evaluated from the streamed layout's device-facing arrays equals the oracle's CSR
    product of the natural-order stencil, for level-1 partitions from the pinned mt-metis binary
    (weighted brick graph) and for scattered owners (many neighbours per rank).
"""
import ctypes as C

import numpy as np
import pytest

from ehyb_spmv_gpu_b200 import _lib as L
from ehyb_spmv_gpu_b200 import api
from ehyb_spmv_gpu_b200 import multigpu as mg
from ehyb_spmv_gpu_b200._lib import check
from oracle import oracle as O
from tests import util


def _view(lib, lay):
    v = api.LayoutView()
    check(lib, lib.ehyb_layout_get(lay, C.byref(v)), "ehyb_layout_get")
    return v


def _raw(lib, lay):
    o = api.Layout.__new__(api.Layout)
    o.lib, o.h, o.v = lib, None, _view(lib, lay)
    return o.raw(), o.v


def _assert_same_layout(lib, a, b):
    ra, va = _raw(lib, a)
    rb, vb = _raw(lib, b)
    for k in ("n", "ncols", "nnz", "nParts", "W", "nSlices", "blobBytes", "nOverflow", "cacheTotal", "cacheMax", "nnzEll",
              "nnzRemInSlice", "nnzOverflow", "padEll", "padRem", "nLongRows", "algBytes", "formatBytes", "haloInOverflow"):
        assert getattr(va, k) == getattr(vb, k), (k, getattr(va, k), getattr(vb, k))
    for k in ra:
        assert np.array_equal(ra[k], rb[k]), k


@pytest.mark.parametrize("chunk", [1, 3, 7])
@pytest.mark.parametrize("fill,cache_cap", [(0.0, 0), (0.5, 0), (0.0, 32), (1.0, -1)])
def test_builder_equals_one_shot(lib, chunk, fill, cache_cap):
    """random sparse matrix with rows of very different lengths, 13 partitions of uneven size"""
    rng = np.random.default_rng(5)
    n, P, W = 3000, 13, 320
    cuts = np.sort(rng.choice(np.arange(1, n), P - 1, replace=False))
    pb = np.concatenate([[0], cuts, [n]]).astype(np.int32)
    rows = []
    for r in range(n):
        k = int(rng.integers(0, 40)) if r % 97 else 700   # a few long rows (overflow candidates)
        near = rng.integers(max(0, r - 150), min(n, r + 150), k)
        far = rng.integers(0, n, k // 4 + 1)
        rows.append(np.unique(np.concatenate([near, far, [r]])).astype(np.int32))
    rowPtr = np.concatenate([[0], np.cumsum([len(c) for c in rows])]).astype(np.int64)
    col = np.concatenate(rows)
    val = rng.uniform(-1, 1, len(col))
    o = L.LayoutOpts()
    o.W = W; o.ctasPerPart = 1; o.er_fill = fill; o.cache_cap = cache_cap; o.min_coverage = -1.0
    one = C.c_void_p()
    check(lib, lib.ehyb_layout_build_csr(C.c_int64(n), rowPtr.ctypes.data_as(L.c_i64_p), col.ctypes.data_as(C.POINTER(C.c_int32)),
                                         val.ctypes.data_as(L.c_dbl_p), P, pb.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(o), C.byref(one)),
          "ehyb_layout_build_csr")
    B = C.c_void_p()
    check(lib, lib.ehyb_layout_builder_begin(C.c_int64(n), C.byref(o), C.byref(B)), "ehyb_layout_builder_begin")
    for p0 in range(0, P, chunk):
        p1 = min(P, p0 + chunk)
        r0, r1 = int(pb[p0]), int(pb[p1])
        rp = (rowPtr[r0:r1 + 1] - rowPtr[r0]).astype(np.int64)
        cc = np.ascontiguousarray(col[rowPtr[r0]:rowPtr[r1]]); vv = np.ascontiguousarray(val[rowPtr[r0]:rowPtr[r1]])
        pbc = np.ascontiguousarray(pb[p0:p1 + 1])
        check(lib, lib.ehyb_layout_builder_add(B, p1 - p0, pbc.ctypes.data_as(C.POINTER(C.c_int32)), rp.ctypes.data_as(L.c_i64_p),
                                               cc.ctypes.data_as(C.POINTER(C.c_int32)), vv.ctypes.data_as(L.c_dbl_p)), "ehyb_layout_builder_add")
    st = C.c_void_p()
    check(lib, lib.ehyb_layout_builder_finish(B, C.byref(st)), "ehyb_layout_builder_finish")
    _assert_same_layout(lib, one, st)
    # and the product from the streamed arrays
    x = rng.uniform(-1, 1, n)
    raw, _ = _raw(lib, st)
    y = util.layout_spmv(raw, x)
    y_ref = np.zeros(n)
    np.add.at(y_ref, np.repeat(np.arange(n), np.diff(rowPtr)), val * x[col])
    assert np.allclose(y, y_ref, rtol=0, atol=1e-12)
    lib.ehyb_layout_free(one); lib.ehyb_layout_free(st)


def _owners(kind, grid, brick, world):
    if kind == "metis":
        return mg.GridDecomp.level1_metis(grid, brick, world)
    if kind == "runs":
        return None
    nb = int(np.prod([-(-g // b) for g, b in zip(grid, brick)]))
    return np.random.default_rng(3).integers(0, world, nb).astype(np.uint32)   # scattered: every rank neighbours every other


@pytest.mark.parametrize("grid,brick,world,owners,exchange", [
    ((12, 10, 9), (4, 4, 4), 3, "scatter", "p2p"),     # clipped bricks at the domain boundary
    ((16, 16, 16), (4, 4, 8), 2, "runs", "nccl"),
    ((24, 24, 24), (6, 6, 6), 4, "metis", "p2p"),
])
def test_grid_stream_equals_general_path(lib, orc, grid, brick, world, owners, exchange):
    own = _owners(owners, grid, brick, world)
    dec = mg.GridDecomp(grid, brick, world, own)
    N = int(np.prod(grid))
    assert int(dec.rowStarts[-1]) == N
    # reference product of the natural-order stencil (oracle CSR), x a function of the natural index
    n_, li, lj, lv = O.gen_stencil27_lower(*grid)
    mo = orc.read_sym(n_, li, lj, lv)
    xg = mg.x_of_global(np.arange(N))
    yg = orc.csr_spmv(mo["rowIdx"], mo["J"], mo["V"], xg)
    ag = orc.csr_abs_spmv(mo["rowIdx"], mo["J"], mo["V"], xg)
    peers = []
    for rank in range(world):
        blk = mg.GridBlock(dec, rank, er_fill=0.0, exchange=exchange, chunk_bricks=3)
        rowPtr, col, val, pv = dec.rows(rank)
        gen = mg.DistributedBlock(rank, world, dec.rowStarts, rowPtr, col, val)
        assert np.array_equal(gen.haloGlobal, blk.haloGlobal) and np.array_equal(gen.recvCount, blk.recvCount)
        nat = blk.natural_ids()
        assert np.array_equal(np.sort(nat), np.sort(dec.natural_ids(np.arange(dec.rowStarts[rank], dec.rowStarts[rank + 1]))))
        # both paths need a send list before their layouts are final: "everybody asks for nothing"
        empty = [[np.zeros(0, np.int64)] * world for _ in range(world)]
        gen.set_send(empty)
        blk.set_send(empty)
        nParts = int(pv.max()) + 1
        W = -(-int(np.prod(brick)) // 64) * 64
        gen.finish(nParts, W, 1, pv, er_fill=0.0, exchange=exchange)
        # the level-2 permutation: closed form (streamed) vs counted + sorted (general path)
        perm = gen.coo["reorderList"]
        nat_gen = np.empty(blk.n, np.int64)
        nat_gen[perm] = dec.natural_ids(np.arange(dec.rowStarts[rank], dec.rowStarts[rank + 1]))
        assert np.array_equal(nat, nat_gen)
        _assert_same_layout(lib, gen.layout, blk.layout)
        # the product from the streamed layout's device-facing arrays, on [x_local | halo]
        raw, v = _raw(lib, blk.layout)
        x_ext = np.concatenate([xg[nat], xg[blk.halo_natural_ids()]])
        y = util.layout_spmv(raw, x_ext)
        assert np.all(np.abs(y - yg[nat]) <= 1e-12 * ag[nat])
        assert v.nnz == int(np.diff(rowPtr).sum())
        peers.append(int(np.count_nonzero(blk.recvCount)))
        gen.free(); blk.free()
    if owners == "scatter":
        assert min(peers) == world - 1
    dec.free()


def test_brick_graph_weights():
    """edge weight = matrix entries between two bricks, vertex weight = cells: checked against a
    direct count on the natural-order stencil"""
    grid, brick = (9, 8, 7), (4, 3, 5)
    xa, ad, vw, aw = mg.GridDecomp.brick_graph(grid, brick)
    nbx, nby, nbz = (-(-g // b) for g, b in zip(grid, brick))
    z, y, x = np.meshgrid(np.arange(grid[2]), np.arange(grid[1]), np.arange(grid[0]), indexing="ij")
    bid = ((z // brick[2]) * nby + y // brick[1]) * nbx + x // brick[0]
    assert np.array_equal(vw, np.bincount(bid.ravel(), minlength=nbx * nby * nbz))
    cnt = {}
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = (z + dz >= 0) & (z + dz < grid[2]) & (y + dy >= 0) & (y + dy < grid[1]) & (x + dx >= 0) & (x + dx < grid[0])
                a = bid[ok]
                b = bid[(z + dz)[ok], (y + dy)[ok], (x + dx)[ok]]
                for u, w in zip(a[a != b], b[a != b]):
                    cnt[(int(u), int(w))] = cnt.get((int(u), int(w)), 0) + 1
    got = {(b, int(ad[e])): int(aw[e]) for b in range(len(vw)) for e in range(xa[b], xa[b + 1])}
    assert got == cnt
