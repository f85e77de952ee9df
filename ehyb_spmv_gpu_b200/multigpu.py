"""Multi-GPU front end: one process per GPU (torchrun), rows distributed in contiguous blocks.

torch.distributed is plumbing only: it carries the set-up exchanges (who needs which x entries,
the CUDA IPC handles of the halo buffers or the NCCL unique id) and the barrier / max-over-ranks
around timed regions.  The per-product data path is C/CUDA inside libehyb.so
(csrc/cuda/ehyb_device.cu "multi-GPU", csrc/cuda/ehyb_kernels.cuh PeerArgs), reached through
include/ehyb.h ehyb_mg_*:

  exchange "p2p"  (default) the main kernel stores the x entries its neighbours need into their
                  halo buffers over NVLink and reads its own halo columns through the
                  shared-memory remainder cache: one launch per product, no collective call;
  exchange "nccl" (baseline) pack kernel + grouped ncclSend/ncclRecv overlapped with the main
                  kernel, halo entries in the overflow list (a second launch).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np

from . import _lib as L
from . import api
from ._lib import MatrixCOO, check


def gen_stencil27_rows(nx, ny, nz, z0, z1):
    lib = L.load()
    rp = L.c_i64_p(); col = L.c_i64_p(); val = L.c_dbl_p()
    check(lib, lib.ehyb_gen_stencil27_rows(nx, ny, C.c_int64(nz), C.c_int64(z0), C.c_int64(z1), C.byref(rp), C.byref(col),
                                           C.byref(val)), "ehyb_gen_stencil27_rows")
    n = (z1 - z0) * nx * ny
    rowPtr = api._np(rp, n + 1, np.int64)
    nnz = int(rowPtr[-1])
    out = rowPtr, api._np(col, nnz, np.int64), api._np(val, nnz, np.float64)
    for p in (rp, col, val):
        lib.ehyb_free_host(p)
    return out


class DistributedBlock:
    """A rank's block of rows (ehyb_mg_local)."""

    def __init__(self, rank, world, rowStarts, rowPtr, colGlobal, val):
        self.lib = L.load()
        self.rank, self.world = rank, world
        self.rowStarts = np.ascontiguousarray(rowStarts, np.int64)
        self.n = int(self.rowStarts[rank + 1] - self.rowStarts[rank])
        rowPtr = np.ascontiguousarray(rowPtr, np.int64)
        colGlobal = np.ascontiguousarray(colGlobal, np.int64)
        val = np.ascontiguousarray(val, np.float64)
        self.h = C.c_void_p()
        check(self.lib, self.lib.ehyb_mg_local_build(rank, world, self.rowStarts.ctypes.data_as(L.c_i64_p),
                                                     rowPtr.ctypes.data_as(L.c_i64_p), colGlobal.ctypes.data_as(L.c_i64_p),
                                                     val.ctypes.data_as(L.c_dbl_p), C.byref(self.h)), "ehyb_mg_local_build")
        nh = C.c_int64(); hg = L.c_i64_p(); rc = L.c_i64_p()
        check(self.lib, self.lib.ehyb_mg_local_halo(self.h, C.byref(nh), C.byref(hg), C.byref(rc)), "ehyb_mg_local_halo")
        self.nHalo = nh.value
        self.haloGlobal = api._np(hg, self.nHalo, np.int64)
        self.recvCount = api._np(rc, world, np.int64)
        self.session = None

    def needs(self):
        """need[g] = global rows this rank receives from rank g (sorted)."""
        off = np.concatenate([[0], np.cumsum(self.recvCount)])
        return [self.haloGlobal[off[g]:off[g + 1]] for g in range(self.world)]

    def set_send(self, all_needs):
        """all_needs[r][g] as gathered from every rank r; this rank sends all_needs[r][self.rank] to r."""
        send = [np.ascontiguousarray(all_needs[r][self.rank], np.int64) for r in range(self.world)]
        cnt = np.array([len(s) for s in send], np.int64)
        cat = np.concatenate(send) if cnt.sum() else np.zeros(1, np.int64)
        check(self.lib, self.lib.ehyb_mg_local_set_send(self.h, cnt.ctypes.data_as(L.c_i64_p), cat.ctypes.data_as(L.c_i64_p)),
              "ehyb_mg_local_set_send")
        self.sendCount = cnt

    def exchange_lists(self, dist):
        gathered = [None] * self.world
        dist.all_gather_object(gathered, self.needs())
        self.set_send(gathered)

    def local_graph(self):
        xadj = L.c_u32_p(); adj = L.c_u32_p()
        check(self.lib, self.lib.ehyb_mg_local_graph(self.h, C.byref(xadj), C.byref(adj)), "ehyb_mg_local_graph")
        xa = api._np(xadj, self.n + 1, np.uint32)
        ad = api._np(adj, int(xa[-1]), np.uint32)
        self.lib.ehyb_free_host(xadj); self.lib.ehyb_free_host(adj)
        return xa, ad

    def finish(self, nParts, W, kpp=1, partVec=None, er_fill=-1.0, exchange="p2p"):
        pv = None
        if partVec is not None:
            pv = np.ascontiguousarray(partVec, np.uint32)
        self.exchange = exchange
        check(self.lib, self.lib.ehyb_mg_local_finish(self.h, nParts, W, kpp, pv.ctypes.data_as(L.c_u32_p) if pv is not None else None,
                                                      C.c_double(er_fill), EXCHANGES[exchange]), "ehyb_mg_local_finish")
        coo = C.POINTER(MatrixCOO)(); lay = C.c_void_p(); ns = C.c_int64(); si = C.POINTER(C.c_int32)(); sc = L.c_i64_p()
        check(self.lib, self.lib.ehyb_mg_local_view(self.h, C.byref(coo), C.byref(lay), C.byref(ns), C.byref(si), C.byref(sc)),
              "ehyb_mg_local_view")
        c = coo.contents
        self.coo = dict(n=c.dimension, nnz=c.totalNum, rowIdx=api._np(c.rowIdx, c.dimension + 1, np.int32),
                        J=api._np(c.J, c.totalNum, np.int32), V=api._np(c.V, c.totalNum, np.float64),
                        reorderList=api._np(c.reorderList, c.dimension, np.int32),
                        partBoundary=api._np(c.partBoundary, nParts + 1, np.int32))
        self.sendIdx = api._np(si, ns.value, np.int32)
        v = api.LayoutView()
        check(self.lib, self.lib.ehyb_layout_get(lay, C.byref(v)), "ehyb_layout_get")
        self.stats = {k: getattr(v, k) for k in ("n", "ncols", "nnz", "nParts", "W", "nSlices", "nOverflow", "nnzEll",
                                                 "nnzRemInSlice", "nnzOverflow", "algBytes", "formatBytes", "cacheMax",
                                                 "haloInOverflow")}
        self.layout = lay

    def recv_offsets_on_peers(self, all_recv_counts):
        """all_recv_counts[g] = rank g's recvCount array.  Entry g of the result: where this rank's
        entries start in rank g's halo list (the list is grouped by owner in rank order)."""
        return np.array([int(np.sum(np.asarray(all_recv_counts[g])[:self.rank])) for g in range(self.world)], np.int64)

    # ---- device ------------------------------------------------------------------------
    def create_session(self, device, unique_id: bytes):
        """NCCL exchange (collective: joins the communicator)."""
        self.session = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        check(self.lib, self.lib.ehyb_mg_session_create(self.h, self.rank, self.world, device, buf, C.byref(self.session)),
              "ehyb_mg_session_create")
        self.handle = C.c_void_p()
        check(self.lib, self.lib.ehyb_mg_session_handle(self.session, C.byref(self.handle)), "ehyb_mg_session_handle")

    def create_session_p2p(self, device, dist):
        """Peer-memory exchange (collective over `dist`: IPC handles and halo offsets are gathered)."""
        self.session = C.c_void_p()
        check(self.lib, self.lib.ehyb_mg_session_create_p2p(self.h, self.rank, self.world, device, C.byref(self.session)),
              "ehyb_mg_session_create_p2p")
        self.handle = C.c_void_p()
        check(self.lib, self.lib.ehyb_mg_session_handle(self.session, C.byref(self.handle)), "ehyb_mg_session_handle")
        blob = C.create_string_buffer(P2P_BLOB_BYTES)
        check(self.lib, self.lib.ehyb_mg_p2p_export(self.session, blob), "ehyb_mg_p2p_export")
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (blob.raw, [int(c) for c in self.recvCount]))
        blobs = C.create_string_buffer(b"".join(g[0] for g in gathered), P2P_BLOB_BYTES * self.world)
        off = self.recv_offsets_on_peers([g[1] for g in gathered])
        check(self.lib, self.lib.ehyb_mg_p2p_connect(self.session, blobs, off.ctypes.data_as(L.c_i64_p)), "ehyb_mg_p2p_connect")
        dist.barrier()  # every rank has mapped its neighbours before anybody pushes

    def launches_per_spmv(self):
        return int(self.lib.ehyb_mg_launches_per_spmv(self.session))

    def kernel_name(self):
        return self.lib.ehyb_session_kernel(self.handle).decode()

    def timed_out(self):
        t = C.c_int()
        check(self.lib, self.lib.ehyb_mg_status(self.session, C.byref(t)), "ehyb_mg_status")
        return bool(t.value)

    def set_x(self, x_local_perm):
        xe = np.zeros(self.n + self.nHalo, np.float64)
        xe[:self.n] = x_local_perm
        check(self.lib, self.lib.ehyb_set_x(self.handle, xe.ctypes.data_as(L.c_dbl_p)), "ehyb_set_x")

    def spmv(self):
        xd = C.c_void_p(); yd = C.c_void_p()
        check(self.lib, self.lib.ehyb_session_vectors(self.handle, C.byref(xd), C.byref(yd)), "ehyb_session_vectors")
        check(self.lib, self.lib.ehyb_mg_spmv(self.session, xd, yd), "ehyb_mg_spmv")
        check(self.lib, self.lib.ehyb_sync(self.handle), "ehyb_sync")

    def get_y(self):
        y = np.empty(self.n, np.float64)
        check(self.lib, self.lib.ehyb_get_y(self.handle, y.ctypes.data_as(L.c_dbl_p)), "ehyb_get_y")
        return y

    def time_spmv(self, warmup, iters):
        ms = C.c_float()
        check(self.lib, self.lib.ehyb_mg_time_spmv(self.session, warmup, iters, C.byref(ms)), "ehyb_mg_time_spmv")
        return ms.value

    def allreduce_sum(self, vals_d, count):
        """ehyb_mg_allreduce_sum: count <= 4 doubles at the device address vals_d, summed over the ranks in place
        (asynchronous, on the session stream; collective)."""
        check(self.lib, self.lib.ehyb_mg_allreduce_sum(self.session, C.c_void_p(vals_d), int(count)), "ehyb_mg_allreduce_sum")

    def pcg_solve(self, b_local, diag_local=None, max_iters=1000, rtol=1e-10, check_every=8):
        """ehyb_mg_pcg_solve: distributed (Jacobi-)PCG; this rank's rows of b, diag and x in the block's
        permuted numbering.  Collective.  Returns (x_local, info)."""
        b = np.ascontiguousarray(b_local, np.float64)
        x = np.empty_like(b)
        d = np.ascontiguousarray(diag_local, np.float64) if diag_local is not None else None
        o = L.PcgOpts(int(max_iters), float(rtol), int(check_every))
        r = L.PcgResult()
        check(self.lib, self.lib.ehyb_mg_pcg_solve(self.session, d.ctypes.data_as(L.c_dbl_p) if d is not None else None,
                                                   b.ctypes.data_as(L.c_dbl_p), x.ctypes.data_as(L.c_dbl_p), C.byref(o), C.byref(r)),
              "ehyb_mg_pcg_solve")
        return x, dict(iters=r.iters, converged=bool(r.converged), rel_residual=r.rel_residual,
                       true_rel_residual=r.true_rel_residual, ms=r.ms)

    def free(self):
        if self.session:
            self.lib.ehyb_mg_session_free(self.session)
            self.session = None
        if self.h:
            self.lib.ehyb_mg_local_free(self.h)
            self.h = C.c_void_p()


class GridDecomp:
    """27-point stencil grid cut into bricks, bricks assigned to ranks (ehyb_grid_decomp)."""

    def __init__(self, grid, brick, world, owner=None):
        self.lib = L.load()
        self.grid, self.brick, self.world = tuple(grid), tuple(brick), world
        self.h = C.c_void_p()
        ow = np.ascontiguousarray(owner, np.uint32) if owner is not None else None
        check(self.lib, self.lib.ehyb_grid_decomp_create(*self.grid, *self.brick, world, ow.ctypes.data_as(L.c_u32_p) if ow is not None else None,
                                                         C.byref(self.h)), "ehyb_grid_decomp_create")
        nb = C.c_int64(); rs = L.c_i64_p(); own = C.POINTER(C.c_int32)()
        check(self.lib, self.lib.ehyb_grid_decomp_info(self.h, C.byref(nb), C.byref(rs), C.byref(own)), "ehyb_grid_decomp_info")
        self.nBricks = nb.value
        self.rowStarts = api._np(rs, world + 1, np.int64)
        self.owner = api._np(own, self.nBricks, np.int32)

    @staticmethod
    def brick_graph(grid, brick):
        lib = L.load()
        nb = C.c_int64(); xadj = L.c_u32_p(); adj = L.c_u32_p(); vw = C.POINTER(C.c_int32)(); aw = C.POINTER(C.c_int32)()
        check(lib, lib.ehyb_grid_brick_graph(*grid, *brick, C.byref(nb), C.byref(xadj), C.byref(adj), C.byref(vw), C.byref(aw)),
              "ehyb_grid_brick_graph")
        xa = api._np(xadj, nb.value + 1, np.uint32)
        out = xa, api._np(adj, int(xa[-1]), np.uint32), api._np(vw, nb.value, np.int32), api._np(aw, int(xa[-1]), np.int32)
        for q in (xadj, adj, vw, aw):
            lib.ehyb_free_host(q)
        return out

    @staticmethod
    def level1_metis(grid, brick, world, nthreads=1, ubvec=1.001):
        """owner[brick] = mt-metis k = world partition of the weighted brick graph (the reference's
        call, reordering.c:270-293, on the coarsened graph)."""
        lib = L.load()
        xa, ad, vw, aw = GridDecomp.brick_graph(grid, brick)
        where = np.zeros(len(vw), np.uint32)
        check(lib, lib.ehyb_partition_graph_weighted(C.c_uint32(len(vw)), xa.ctypes.data_as(L.c_u32_p), ad.ctypes.data_as(L.c_u32_p),
                                                     vw.ctypes.data_as(C.POINTER(C.c_int32)), aw.ctypes.data_as(C.POINTER(C.c_int32)),
                                                     C.c_uint32(world), C.c_uint32(nthreads), C.c_float(ubvec),
                                                     where.ctypes.data_as(L.c_u32_p)), "ehyb_partition_graph_weighted")
        return where

    def rows(self, rank):
        """All rows of a rank (level-1 order and column ids) + partVec = brick ordinal: the general path's input."""
        rp = L.c_i64_p(); col = L.c_i64_p(); val = L.c_dbl_p(); pv = L.c_u32_p()
        check(self.lib, self.lib.ehyb_grid_rows(self.h, rank, C.byref(rp), C.byref(col), C.byref(val), C.byref(pv)), "ehyb_grid_rows")
        n = int(self.rowStarts[rank + 1] - self.rowStarts[rank])
        rowPtr = api._np(rp, n + 1, np.int64)
        out = rowPtr, api._np(col, int(rowPtr[-1]), np.int64), api._np(val, int(rowPtr[-1]), np.float64), api._np(pv, n, np.uint32)
        for q in (rp, col, val, pv):
            self.lib.ehyb_free_host(q)
        return out

    def natural_ids(self, level1):
        level1 = np.ascontiguousarray(level1, np.int64)
        out = np.empty(len(level1), np.int64)
        check(self.lib, self.lib.ehyb_grid_natural_ids(self.h, C.c_int64(len(level1)), level1.ctypes.data_as(L.c_i64_p),
                                                       out.ctypes.data_as(L.c_i64_p)), "ehyb_grid_natural_ids")
        return out

    def free(self):
        if self.h:
            self.lib.ehyb_grid_decomp_free(self.h)
            self.h = C.c_void_p()


class GridBlock(DistributedBlock):
    """A rank's block of a brick-decomposed stencil grid, streamed into the tuned layout
    (ehyb_mg_grid_build): no matrixCOO exists; x and y live in permuted local order and
    natural_ids() tells which grid cell every entry belongs to."""

    def __init__(self, decomp: GridDecomp, rank, er_fill=0.0, exchange="p2p", chunk_bricks=0):
        self.lib = L.load()
        self.decomp = decomp
        self.rank, self.world = rank, decomp.world
        self.rowStarts = decomp.rowStarts.copy()
        self.n = int(self.rowStarts[rank + 1] - self.rowStarts[rank])
        self.exchange = exchange
        self.h = C.c_void_p()
        check(self.lib, self.lib.ehyb_mg_grid_build(decomp.h, rank, C.c_double(er_fill), EXCHANGES[exchange], chunk_bricks, C.byref(self.h)),
              "ehyb_mg_grid_build")
        nh = C.c_int64(); hg = L.c_i64_p(); rc = L.c_i64_p()
        check(self.lib, self.lib.ehyb_mg_local_halo(self.h, C.byref(nh), C.byref(hg), C.byref(rc)), "ehyb_mg_local_halo")
        self.nHalo = nh.value
        self.haloGlobal = api._np(hg, self.nHalo, np.int64)
        self.recvCount = api._np(rc, self.world, np.int64)
        self.session = None
        self.coo = None
        self._view()

    def _view(self):
        coo = C.POINTER(MatrixCOO)(); lay = C.c_void_p(); ns = C.c_int64(); si = C.POINTER(C.c_int32)(); sc = L.c_i64_p()
        check(self.lib, self.lib.ehyb_mg_local_view(self.h, C.byref(coo), C.byref(lay), C.byref(ns), C.byref(si), C.byref(sc)),
              "ehyb_mg_local_view")
        self.sendIdx = api._np(si, ns.value, np.int32) if ns.value else np.zeros(0, np.int32)
        v = api.LayoutView()
        check(self.lib, self.lib.ehyb_layout_get(lay, C.byref(v)), "ehyb_layout_get")
        self.stats = {k: getattr(v, k) for k in ("n", "ncols", "nnz", "nParts", "W", "nSlices", "nOverflow", "nnzEll",
                                                 "nnzRemInSlice", "nnzOverflow", "algBytes", "formatBytes", "cacheMax",
                                                 "haloInOverflow")}
        self.layout = lay

    def set_send(self, all_needs):
        super().set_send(all_needs)
        self._view()  # the send list in permuted local numbering exists now

    def natural_ids(self):
        out = np.empty(self.n, np.int64)
        check(self.lib, self.lib.ehyb_mg_local_natural_ids(self.h, out.ctypes.data_as(L.c_i64_p)), "ehyb_mg_local_natural_ids")
        return out

    def halo_natural_ids(self):
        return self.decomp.natural_ids(self.haloGlobal)


def setup_grid(rank, world, grid, brick, dist, level1="metis", exchange="p2p", er_fill=0.0):
    """Rank's block of the 27-point stencil on `grid` cut into `brick`s: level 1 = mt-metis on the
    brick graph (rank 0 runs it, every rank gets the owner vector) or contiguous runs of bricks."""
    owner = None
    if level1 == "metis" and world > 1:
        box = [GridDecomp.level1_metis(grid, brick, world) if rank == 0 else None]
        if dist is not None:
            dist.broadcast_object_list(box, src=0)
        owner = box[0]
    dec = GridDecomp(grid, brick, world, owner)
    blk = GridBlock(dec, rank, er_fill=er_fill, exchange=exchange)
    if world > 1:
        blk.exchange_lists(dist)
    else:
        blk.set_send([[np.zeros(0, np.int64)]])
    return blk, dec


EXCHANGES = {"nccl": 0, "p2p": 1}  # EHYB_MG_NCCL, EHYB_MG_P2P
P2P_BLOB_BYTES = 128               # EHYB_MG_P2P_BLOB_BYTES


def p2p_supported(device, world) -> bool:
    lib = L.load()
    ok = C.c_int()
    check(lib, lib.ehyb_mg_p2p_supported(device, world, C.byref(ok)), "ehyb_mg_p2p_supported")
    return bool(ok.value)


def unique_id() -> bytes:
    lib = L.load()
    buf = C.create_string_buffer(128)
    check(lib, lib.ehyb_mg_unique_id(buf), "ehyb_mg_unique_id")
    return buf.raw


def x_of_global(idx):
    """Deterministic x as a function of the global row index (every rank can evaluate any
    entry, which is what the parity check of a distributed product needs)."""
    idx = np.asarray(idx, np.uint64)
    with np.errstate(over="ignore"):
        z = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x632BE59BD9B4E019)
        z ^= z >> np.uint64(29)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(32)
    return ((z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53)) - 0.5) * 0.2


def setup_slab(rank, world, grid, dist, partition="metis", exchange="p2p"):
    """Rank's z-slab of the 27-point stencil on nx x ny x (nz*world)."""
    nx, ny, nzl = grid
    plane = nx * ny
    rowStarts = np.arange(world + 1, dtype=np.int64) * (nzl * plane)
    rowPtr, col, val = gen_stencil27_rows(nx, ny, nzl * world, rank * nzl, (rank + 1) * nzl)
    blk = DistributedBlock(rank, world, rowStarts, rowPtr, col, val)
    del rowPtr, col, val
    blk.exchange_lists(dist)
    try:
        dev = api.device_query(int(os.environ.get("LOCAL_RANK", "0")))
    except Exception:
        dev = api.device_info_b200()
    # partitions sized for the persistent kernel (which carries the halo exchange) unless
    # EHYB_MG_PLAN=staged asks for the staged kernel's two-per-SM plan
    pl = api.plan(blk.n, dev, kernel=api.KERNEL_STAGED if os.environ.get("EHYB_MG_PLAN") == "staged" else api.KERNEL_PERSISTENT)
    pv = None
    if partition == "metis":
        xa, ad = blk.local_graph()
        pv = np.zeros(blk.n, np.uint32)
        check(blk.lib, blk.lib.ehyb_partition_graph(C.c_uint32(blk.n), xa.ctypes.data_as(L.c_u32_p), ad.ctypes.data_as(L.c_u32_p),
                                                    C.c_uint32(pl.nParts), C.c_uint32(1), pv.ctypes.data_as(L.c_u32_p)),
              "ehyb_partition_graph")
    blk.finish(pl.nParts, pl.W, pl.ctasPerPart, pv, exchange=exchange)
    return blk, rowStarts
