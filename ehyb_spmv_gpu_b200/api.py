"""Thin Python mirror of the C ABI (include/ehyb.h, spmv.h, convert.h, reordering.h).

Names follow the C entry points; every method is a direct call into libehyb.so.  The host
language of the engine is C (bin/spmv.out is the reference's driver); this module exists so that
the parity tests and bench.py can drive the same entry points with numpy buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import (DeviceInfo, EhybError, LayoutOpts, LayoutView, MatrixCOO, MatrixEHYB, Plan,
                   SessionOpts, check)

GEN_LAPLACE2D, GEN_STENCIL27, GEN_ELASTICITY = 1, 2, 3
KERNEL_DIRECT, KERNEL_STAGED, KERNEL_PERSISTENT = 1, 2, 3


def _np(ptr, count, dtype):
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    ct = np.ctypeslib.as_ctypes_type(np.dtype(dtype))
    addr = C.cast(ptr, C.c_void_p).value
    return np.ctypeslib.as_array((ct * count).from_address(addr)).copy()


def _p(a, typ):
    return a.ctypes.data_as(typ)


def device_info_b200() -> DeviceInfo:
    d = DeviceInfo()
    L.load().ehyb_device_info_b200(C.byref(d))
    return d


def device_query(device: int = 0) -> DeviceInfo:
    lib = L.load()
    d = DeviceInfo()
    check(lib, lib.ehyb_device_query(device, C.byref(d)), "ehyb_device_query")
    return d


def plan(n: int, dev: DeviceInfo | None = None, kernel: int = 2) -> Plan:
    """ehyb_plan_kernel: partition parameters for the staged kernel (default; what the multi-GPU
    blocks use) or, with kernel=KERNEL_PERSISTENT, for the persistent kernel of single-GPU sessions."""
    lib = L.load()
    p = Plan()
    dev = dev or device_info_b200()
    check(lib, lib.ehyb_plan_kernel(n, C.byref(dev), int(kernel), C.byref(p)), "ehyb_plan_kernel")
    return p


def plan_auto(n: int, nnz: int, dev: DeviceInfo | None = None):
    """ehyb_plan_auto: (plan, kernel) for a matrix of n rows and nnz entries - one partition per SM when it is
    L2-resident, the persistent kernel's plan up to ~40 entries per row, the staged plan beyond."""
    lib = L.load()
    p = Plan()
    k = C.c_int()
    dev = dev or device_info_b200()
    check(lib, lib.ehyb_plan_auto(n, C.c_int64(nnz), C.byref(dev), C.byref(p), C.byref(k)), "ehyb_plan_auto")
    return p, k.value


def plan_reference(n: int, symmetric: bool = True) -> Plan:
    lib = L.load()
    p = Plan()
    check(lib, lib.ehyb_plan_reference(n, 1 if symmetric else 0, C.byref(p)), "ehyb_plan_reference")
    return p


def x_reference(n: int) -> np.ndarray:
    x = np.empty(n, np.float64)
    L.load().ehyb_x_reference(n, _p(x, L.c_dbl_p))
    return x


def gen_lower(kind: int, nx: int, ny: int, nz: int = 1):
    lib = L.load()
    n = C.c_int(); cnt = C.c_int64()
    li = L.c_int_p(); lj = L.c_int_p(); lv = L.c_dbl_p()
    check(lib, lib.ehyb_gen_lower(kind, nx, ny, nz, C.byref(n), C.byref(cnt), C.byref(li), C.byref(lj), C.byref(lv)),
          "ehyb_gen_lower")
    out = n.value, _np(li, cnt.value, np.int32), _np(lj, cnt.value, np.int32), _np(lv, cnt.value, np.float64)
    for p in (li, lj, lv):
        lib.ehyb_free_host(p)
    return out


def gen_rmat(scale: int, edge_factor: int = 16, seed: int = 1, add_diagonal: bool = False):
    lib = L.load()
    n = C.c_int(); cnt = C.c_int64()
    fi = L.c_int_p(); fj = L.c_int_p(); fv = L.c_dbl_p()
    check(lib, lib.ehyb_gen_rmat(scale, edge_factor, C.c_uint64(seed), 1 if add_diagonal else 0, C.byref(n),
                                 C.byref(cnt), C.byref(fi), C.byref(fj), C.byref(fv)), "ehyb_gen_rmat")
    out = n.value, _np(fi, cnt.value, np.int32), _np(fj, cnt.value, np.int32), _np(fv, cnt.value, np.float64)
    for p in (fi, fj, fv):
        lib.ehyb_free_host(p)
    return out


class CooMatrix:
    """Owner of a C matrixCOO (include/spmv.h).  Arrays live in C memory (the reorder stage
    frees and replaces I/J/V exactly like the reference, reordering.c:363-369)."""

    def __init__(self):
        self.lib = L.load()
        self.c = MatrixCOO()
        self.symmetric = True
        self.y_golden = None
        self._alive = False

    # -- constructors ---------------------------------------------------------------------
    @classmethod
    def from_lower(cls, n, li, lj, lv, x=None):
        m = cls()
        li = np.ascontiguousarray(li, np.int32); lj = np.ascontiguousarray(lj, np.int32)
        lv = np.ascontiguousarray(lv, np.float64)
        yg = np.zeros(n, np.float64) if x is not None else None
        check(m.lib, m.lib.ehyb_coo_from_lower(n, C.c_int64(len(li)), _p(li, L.c_int_p), _p(lj, L.c_int_p),
                                                _p(lv, L.c_dbl_p), C.byref(m.c),
                                                _p(x, L.c_dbl_p) if x is not None else None,
                                                _p(yg, L.c_dbl_p) if yg is not None else None), "ehyb_coo_from_lower")
        m.symmetric, m.y_golden, m._alive = True, yg, True
        return m

    @classmethod
    def from_general(cls, n, fi, fj, fv, x=None):
        m = cls()
        fi = np.ascontiguousarray(fi, np.int32); fj = np.ascontiguousarray(fj, np.int32)
        fv = np.ascontiguousarray(fv, np.float64)
        yg = np.zeros(n, np.float64) if x is not None else None
        check(m.lib, m.lib.ehyb_coo_from_general(n, C.c_int64(len(fi)), _p(fi, L.c_int_p), _p(fj, L.c_int_p),
                                                  _p(fv, L.c_dbl_p), C.byref(m.c),
                                                  _p(x, L.c_dbl_p) if x is not None else None,
                                                  _p(yg, L.c_dbl_p) if yg is not None else None), "ehyb_coo_from_general")
        m.symmetric, m.y_golden, m._alive = False, yg, True
        return m

    @classmethod
    def generate(cls, kind, nx, ny, nz=1, x=None):
        n, li, lj, lv = gen_lower(kind, nx, ny, nz)
        return cls.from_lower(n, li, lj, lv, x)

    @classmethod
    def read_mtx(cls, path):
        m = cls()
        sym = C.c_int(); x = L.c_dbl_p(); y = L.c_dbl_p()
        check(m.lib, m.lib.ehyb_read_mtx(str(path).encode(), C.byref(m.c), C.byref(sym), C.byref(x), C.byref(y)),
              "ehyb_read_mtx")
        m.symmetric = bool(sym.value)
        m.x = _np(x, m.c.dimension, np.float64)
        m.y_golden = _np(y, m.c.dimension, np.float64)
        m.lib.ehyb_free_host(x); m.lib.ehyb_free_host(y)
        m._alive = True
        return m

    # -- properties -------------------------------------------------------------------------
    @property
    def n(self):
        return self.c.dimension

    @property
    def nnz(self):
        return self.c.totalNum

    def set_plan(self, nParts, W, kpp=1):
        self.c.nParts = int(nParts)
        self.c.vectorCacheSize = int(W)
        self.c.kernelPerPart = int(max(kpp, 0))

    def arrays(self):
        c, n, nnz = self.c, self.c.dimension, self.c.totalNum
        d = dict(n=n, nnz=nnz, nParts=c.nParts, W=c.vectorCacheSize, kpp=c.kernelPerPart, maxCol=c.maxCol,
                 I=_np(c.I, nnz, np.int32), J=_np(c.J, nnz, np.int32), V=_np(c.V, nnz, np.float64),
                 rowIdx=_np(c.rowIdx, n + 1, np.int32), numInRow=_np(c.numInRow, n, np.int32),
                 numInRow2=_np(c.numInRow2, n, np.int32), reorderList=_np(c.reorderList, n, np.int32),
                 diag=_np(c.diag, n, np.float64), symmetric=self.symmetric)
        if c.nParts > 0 and c.partBoundary:
            d["partBoundary"] = _np(c.partBoundary, c.nParts + 1, np.int32)
        return d

    # -- pipeline stages ----------------------------------------------------------------------
    def build_graph(self):
        xadj = L.c_u32_p(); adj = L.c_u32_p()
        check(self.lib, self.lib.ehyb_build_graph(C.byref(self.c), 1 if self.symmetric else 0, C.byref(xadj),
                                                  C.byref(adj)), "ehyb_build_graph")
        xa = _np(xadj, self.n + 1, np.uint32)
        ad = _np(adj, int(xa[-1]), np.uint32)
        self.lib.ehyb_free_host(xadj); self.lib.ehyb_free_host(adj)
        return xa, ad

    def reorder(self):
        """matrixReorder / matrixReorder_unsym (runs the partitioner)."""
        check(self.lib, self.lib.ehyb_reorder(C.byref(self.c), 1 if self.symmetric else 0), "ehyb_reorder")

    def reorder_with_partition(self, partVec):
        pv = np.ascontiguousarray(partVec, np.uint32)
        assert len(pv) == self.n
        check(self.lib, self.lib.ehyb_reorder_with_partition(C.byref(self.c), _p(pv, L.c_u32_p)),
              "ehyb_reorder_with_partition")

    def vector_reorder(self, v):
        out = np.empty(self.n, np.float64)
        v = np.ascontiguousarray(v, np.float64)
        self.lib.vectorReorder(self.n, _p(v, L.c_dbl_p), _p(out, L.c_dbl_p), self.c.reorderList)
        return out

    def vector_recover(self, vr):
        out = np.empty(self.n, np.float64)
        vr = np.ascontiguousarray(vr, np.float64)
        self.lib.vectorRecover(self.n, _p(vr, L.c_dbl_p), _p(out, L.c_dbl_p), self.c.reorderList)
        return out

    def coo2ehyb(self):
        """COO2EHYB: the reference layout as numpy arrays."""
        e = MatrixEHYB(); sE = C.c_int(); sR = C.c_int()
        self.lib.COO2EHYB(C.byref(self.c), C.byref(e), C.byref(sE), C.byref(sR))
        out = ehyb_struct_to_dict(e, sE.value, sR.value, self.c.dimension)
        self.lib.EHYBfreeHost(C.byref(e))
        return out

    def free(self):
        if self._alive:
            self.lib.ehyb_coo_free(C.byref(self.c))
            self._alive = False

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def ehyb_struct_to_dict(e: MatrixEHYB, sizeELL: int, sizeER: int, n: int):
    P, W = e.nParts, e.vectorCacheSize
    S = W // 32
    nER = e.numOfRowER
    nb = (nER + 31) // 32
    nl = e.nLongVec
    d = dict(n=n, nParts=P, W=W, partBoundary=_np(e.partBoundary, P + 1, np.int32),
             widthVecBlockELL=_np(e.widthVecBlockELL, P * S, np.int16),
             biasVecBlockELL=_np(e.biasVecBlockELL, P * S, np.int32),
             colBlockELL=_np(e.colBlockELL, sizeELL, np.int16), valBlockELL=_np(e.valBlockELL, sizeELL, np.float64),
             numOfRowER=nER, reorderListER=_np(e.reorderListER, n, np.int32), rowVecER=_np(e.rowVecER, nER, np.int32),
             widthVecER=_np(e.widthVecER, nb, np.int16), biasVecER=_np(e.biasVecER, nb, np.int32),
             colER=_np(e.colER, sizeER, np.int32), valER=_np(e.valER, sizeER, np.float64),
             sizeBlockELL=sizeELL, sizeER=sizeER, nLongVec=nl)
    if nl:
        d["longVecBoundary"] = _np(e.longVecBoundary, nl + 1, np.int32)
        d["longRow"] = _np(e.longVecRow, nl, np.int32)
        tot = int(d["longVecBoundary"][-1])
        d["longVecCol"] = _np(e.longVecCol, tot, np.int32)
        d["longVecVal"] = _np(e.longVecVal, tot, np.float64)
    else:
        d["longRow"] = np.zeros(0, np.int32)
    return d


class Layout:
    """ehyb_layout: the Blackwell-tuned layout on the host."""

    def __init__(self, m: CooMatrix, W: int = 0, ctasPerPart: int = 0, er_fill: float = -1.0,
                 long_row_threshold: int = 0, ncols: int = 0, halo_in_overflow: bool = False, cache_cap: int = 0,
                 min_coverage: float = 0.0):
        self.lib = L.load()
        self.h = C.c_void_p()
        o = LayoutOpts(W, ctasPerPart, er_fill, long_row_threshold, ncols, int(halo_in_overflow), cache_cap, min_coverage)
        check(self.lib, self.lib.ehyb_layout_build(C.byref(m.c), C.byref(o), C.byref(self.h)), "ehyb_layout_build")
        self.v = LayoutView()
        check(self.lib, self.lib.ehyb_layout_get(self.h, C.byref(self.v)), "ehyb_layout_get")

    @classmethod
    def _adopt(cls, handle):
        self = cls.__new__(cls)
        self.lib = L.load()
        self.h = handle
        self.v = LayoutView()
        check(self.lib, self.lib.ehyb_layout_get(self.h, C.byref(self.v)), "ehyb_layout_get")
        return self

    def save(self, path):
        """ehyb_layout_save: the layout alone as a binary file."""
        check(self.lib, self.lib.ehyb_layout_save(self.h, str(path).encode()), "ehyb_layout_save")

    @classmethod
    def load(cls, path):
        lib = L.load()
        h = C.c_void_p()
        check(lib, lib.ehyb_layout_load(str(path).encode(), C.byref(h)), "ehyb_layout_load")
        return cls._adopt(h)

    def cache_save(self, path, source_path, symmetric, reorderList, x, y_golden, absAx):
        """ehyb_cache_save: the finished pipeline of `source_path` (layout + permutation + vectors)."""
        rl = np.ascontiguousarray(reorderList, np.int32)
        vs = [np.ascontiguousarray(a, np.float64) for a in (x, y_golden, absAx)]
        check(self.lib, self.lib.ehyb_cache_save(str(path).encode(), str(source_path).encode() if source_path else None, self.h,
                                                 int(symmetric), rl.ctypes.data_as(L.c_int_p), *[a.ctypes.data_as(L.c_dbl_p) for a in vs]),
              "ehyb_cache_save")

    @classmethod
    def cache_load(cls, path, source_path=None, plan=None):
        """ehyb_cache_load -> (Layout, dict(n, symmetric, reorderList, x, y_golden, absAx))."""
        lib = L.load()
        h = C.c_void_p(); n = C.c_int(); sym = C.c_int()
        rl = L.c_int_p(); x = L.c_dbl_p(); y = L.c_dbl_p(); a = L.c_dbl_p()
        check(lib, lib.ehyb_cache_load(str(path).encode(), str(source_path).encode() if source_path else None,
                                       C.byref(plan) if plan is not None else None, C.byref(h), C.byref(n), C.byref(sym),
                                       C.byref(rl), C.byref(x), C.byref(y), C.byref(a)), "ehyb_cache_load")
        out = dict(n=n.value, symmetric=bool(sym.value))
        for k, ptr, dt in (("reorderList", rl, np.int32), ("x", x, np.float64), ("y_golden", y, np.float64), ("absAx", a, np.float64)):
            out[k] = _np(ptr, n.value, dt) if ptr else None
            if ptr:
                lib.ehyb_free_host(ptr)
        return cls._adopt(h), out

    def stats(self):
        v = self.v
        return {k: getattr(v, k) for k in ("n", "ncols", "nnz", "nParts", "W", "ctasPerPart", "nSlices", "blobBytes",
                                           "nOverflow", "nnzEll", "nnzRemInSlice", "nnzOverflow", "padEll", "padRem",
                                           "nLongRows", "algBytes", "formatBytes", "cacheTotal", "cacheMax")}

    def raw(self):
        """Copies of the device-facing arrays (for the independent numpy decoder in tests)."""
        v = self.v
        parts = _np(v.parts, v.nParts * 8, np.int32).reshape(v.nParts, 8)
        sl = np.frombuffer(C.string_at(C.cast(v.slices, C.c_void_p).value, v.nSlices * 8) if v.nSlices else b"",
                           dtype=np.dtype([("off256", "<u4"), ("w", "<u2"), ("wr", "<u2")]))
        blob = np.frombuffer(C.string_at(v.blob, v.blobBytes) if v.blobBytes else b"", dtype=np.uint8)
        return dict(parts=parts, slices=sl, blob=blob, cacheCols=_np(v.cacheCols, v.cacheTotal, np.int32),
                    ovfRow=_np(v.ovfRow, v.nOverflow, np.int32),
                    ovfCol=_np(v.ovfCol, v.nOverflow, np.int32), ovfVal=_np(v.ovfVal, v.nOverflow, np.float64))

    def to_reference(self):
        e = MatrixEHYB(); sE = C.c_int(); sR = C.c_int()
        check(self.lib, self.lib.ehyb_layout_to_reference(self.h, C.byref(e), C.byref(sE), C.byref(sR)),
              "ehyb_layout_to_reference")
        out = ehyb_struct_to_dict(e, sE.value, sR.value, int(self.v.n))
        pb = e.partBoundary
        self.lib.EHYBfreeHost(C.byref(e))
        self.lib.ehyb_free_host(pb)
        return out

    def free(self):
        if self.h:
            self.lib.ehyb_layout_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Session:
    """ehyb_handle: a layout resident on one GPU."""

    def __init__(self, layout: Layout, device: int = 0, threads: int = 0, use_graph: bool = True,
                 l2_persist_x: bool = True, halo_cols: int = 0, kernel: int = 0):
        self.lib = L.load()
        self.h = C.c_void_p()
        self.n = int(layout.v.n)
        self.ncols = int(layout.v.ncols)
        o = SessionOpts()
        self.lib.ehyb_session_opts_default(C.byref(o))
        o.device, o.threads, o.use_graph, o.l2_persist_x, o.halo_cols = device, threads, int(use_graph), int(l2_persist_x), halo_cols
        o.kernel = kernel
        check(self.lib, self.lib.ehyb_upload(layout.h, C.byref(o), C.byref(self.h)), "ehyb_upload")

    def spmv_host(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float64)
        assert len(x) == self.ncols
        y = np.empty(self.n, np.float64)
        check(self.lib, self.lib.ehyb_spmv_host(self.h, _p(x, L.c_dbl_p), _p(y, L.c_dbl_p)), "ehyb_spmv_host")
        return y

    def spmv_host_batch(self, xs, ys):
        k = len(xs)
        xp = (L.c_dbl_p * k)(*[_p(a, L.c_dbl_p) for a in xs])
        yp = (L.c_dbl_p * k)(*[_p(a, L.c_dbl_p) for a in ys])
        check(self.lib, self.lib.ehyb_spmv_host_batch(self.h, xp, yp, k), "ehyb_spmv_host_batch")

    def set_x(self, x):
        x = np.ascontiguousarray(x, np.float64)
        check(self.lib, self.lib.ehyb_set_x(self.h, _p(x, L.c_dbl_p)), "ehyb_set_x")

    def get_y(self):
        y = np.empty(self.n, np.float64)
        check(self.lib, self.lib.ehyb_get_y(self.h, _p(y, L.c_dbl_p)), "ehyb_get_y")
        return y

    def spmv_resident(self):
        """One product of the session's own x into its own y (device resident)."""
        xd = C.c_void_p(); yd = C.c_void_p()
        check(self.lib, self.lib.ehyb_session_vectors(self.h, C.byref(xd), C.byref(yd)), "ehyb_session_vectors")
        check(self.lib, self.lib.ehyb_spmv(self.h, xd, yd), "ehyb_spmv")
        check(self.lib, self.lib.ehyb_sync(self.h), "ehyb_sync")

    def time_spmv(self, warmup: int, iters: int, kernel_only: bool = False):
        ms = C.c_float(); kms = C.c_float()
        check(self.lib, self.lib.ehyb_time_spmv(self.h, warmup, iters, C.byref(ms), C.byref(kms) if kernel_only else None),
              "ehyb_time_spmv")
        return (ms.value, kms.value) if kernel_only else ms.value

    def spmv_dot_supported(self) -> bool:
        return bool(self.lib.ehyb_spmv_dot_supported(self.h))

    def spmv_dot_host(self, x):
        """ehyb_spmv_dot_host: (y, x . y) with the dot product computed inside the product kernel."""
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty(self.n if hasattr(self, "n") else len(x), np.float64)
        d = C.c_double()
        check(self.lib, self.lib.ehyb_spmv_dot_host(self.h, x.ctypes.data_as(L.c_dbl_p), y.ctypes.data_as(L.c_dbl_p), C.byref(d)), "ehyb_spmv_dot_host")
        return y, d.value

    def time_spmv_flushed(self, warmup: int, iters: int, flush_bytes: int = 512 << 20) -> float:
        """Sum (ms) of `iters` products each timed alone behind an L2 flush (cold matrix, x and y)."""
        ms = C.c_float()
        check(self.lib, self.lib.ehyb_time_spmv_flushed(self.h, warmup, iters, C.c_size_t(flush_bytes), C.byref(ms)), "ehyb_time_spmv_flushed")
        return ms.value

    def pcg_solve(self, b, diag=None, max_iters=1000, rtol=1e-10, check_every=8):
        """ehyb_pcg_solve: A x = b (permuted numbering), Jacobi-preconditioned when diag is given.
        Returns (x, dict(iters, converged, rel_residual, true_rel_residual, ms))."""
        b = np.ascontiguousarray(b, np.float64)
        x = np.empty_like(b)
        d = np.ascontiguousarray(diag, np.float64) if diag is not None else None
        o = L.PcgOpts(int(max_iters), float(rtol), int(check_every))
        r = L.PcgResult()
        check(self.lib, self.lib.ehyb_pcg_solve(self.h, d.ctypes.data_as(L.c_dbl_p) if d is not None else None,
                                                b.ctypes.data_as(L.c_dbl_p), x.ctypes.data_as(L.c_dbl_p), C.byref(o), C.byref(r)),
              "ehyb_pcg_solve")
        return x, dict(iters=r.iters, converged=bool(r.converged), rel_residual=r.rel_residual,
                       true_rel_residual=r.true_rel_residual, ms=r.ms)

    def kernel_name(self) -> str:
        return self.lib.ehyb_session_kernel(self.h).decode()

    def launches_per_spmv(self) -> int:
        return int(self.lib.ehyb_launches_per_spmv(self.h))

    def free(self):
        if self.h:
            self.lib.ehyb_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
