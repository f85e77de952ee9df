"""ctypes front end to the checker libraries.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/ehyb_oracle.c).  It wraps

  * oracle/_build/libehyb_oracle.so -- the CPU restatement (ehyb_oracle.c), and
  * oracle/_ref/libehyb_ref.so      -- the unmodified reference host path (ref_shim.cpp),
                                       present when it was built where /root/reference exists.

It also holds the synthetic matrix generators of SURVEY.md section 8(d) in numpy, written
independently of the product's C generators so that the two can be checked against each other.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
ORACLE_SO = HERE / "_build" / "libehyb_oracle.so"
REF_SO = HERE / "_ref" / "libehyb_ref.so"
MTMETIS_BIN = ROOT / "bin" / "ehyb_mtmetis"

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_i16_p = C.POINTER(C.c_int16)
c_u32_p = C.POINTER(C.c_uint32)


def _p(a, typ):
    return a.ctypes.data_as(typ)


def build(force: bool = False) -> None:
    """Compile the checker libraries (oracle always; _ref only where the reference exists)."""
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < (HERE / "ehyb_oracle.c").stat().st_mtime:
        subprocess.check_call(["make", "-C", str(HERE), "oracle"], stdout=subprocess.DEVNULL)
    if Path("/root/reference/convert.c").exists():
        subprocess.check_call(["make", "-C", str(HERE), "ref"], stdout=subprocess.DEVNULL)


# --------------------------------------------------------------------------------------
# struct layouts of the reference (spmv.h:17-63)
# --------------------------------------------------------------------------------------
class MatrixCOO(C.Structure):
    _fields_ = [
        ("totalNum", C.c_int), ("dimension", C.c_int), ("maxCol", C.c_int), ("nParts", C.c_int),
        ("vectorCacheSize", C.c_uint16), ("kernelPerPart", C.c_int16),
        ("rowIdx", c_int_p), ("numInRow", c_int_p), ("numInRow2", c_int_p),
        ("I", c_int_p), ("J", c_int_p), ("V", c_dbl_p), ("diag", c_dbl_p),
        ("partBoundary", c_int_p), ("reorderList", c_int_p),
    ]


class MatrixEHYB(C.Structure):
    _fields_ = [
        ("dimension", C.c_int), ("nParts", C.c_int), ("vectorCacheSize", C.c_int16),
        ("kernelPerPart", C.c_int), ("numOfRowER", C.c_int), ("warpIdxER_d", c_int_p),
        ("reorderList", c_int_p), ("reorderListER", c_int_p),
        ("widthVecBlockELL", c_i16_p), ("biasVecBlockELL", c_int_p),
        ("colBlockELL", c_i16_p), ("valBlockELL", c_dbl_p), ("partBoundary", c_int_p),
        ("widthVecER", c_i16_p), ("rowVecER", c_int_p), ("biasVecER", c_int_p),
        ("colER", c_int_p), ("valER", c_dbl_p), ("outER", c_dbl_p),
        ("nLongVec", C.c_int), ("longVecBoundary", c_int_p), ("longVecRow", c_int_p),
        ("longVecCol", c_int_p), ("longVecVal", c_dbl_p),
    ]


def np_from(ptr, count, dtype):
    """Copy `count` elements out of a C pointer."""
    if count == 0:
        return np.zeros(0, dtype=dtype)
    ct = np.ctypeslib.as_ctypes_type(np.dtype(dtype))
    buf = C.cast(ptr, C.POINTER(ct * count)).contents
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


# --------------------------------------------------------------------------------------
# generators; symmetric ones return the lower-triangle FILE entries
# in file order (0-based), exactly what a .mtx written per the spec contains.
# --------------------------------------------------------------------------------------
def gen_laplace2d_lower(nx: int, ny: int):
    """2-D 5-point Laplacian, diag 4, off -1; per column c: (c,c), (c+1,c), (c+nx,c)."""
    n = nx * ny
    c = np.arange(n, dtype=np.int64)
    has_e = (c % nx) != nx - 1
    has_s = c < n - nx
    cnt = 1 + has_e.astype(np.int64) + has_s.astype(np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=off[1:])
    m = int(off[-1])
    li = np.empty(m, np.int32); lj = np.empty(m, np.int32); lv = np.empty(m, np.float64)
    li[off[:-1]] = c; lj[off[:-1]] = c; lv[off[:-1]] = 4.0
    pe = off[:-1][has_e] + 1
    li[pe] = c[has_e] + 1; lj[pe] = c[has_e]; lv[pe] = -1.0
    ps = off[:-1][has_s] + 1 + has_e[has_s]
    li[ps] = c[has_s] + nx; lj[ps] = c[has_s]; lv[ps] = -1.0
    return n, li, lj, lv


def _stencil27_pairs(nx, ny, nz):
    """(r_node, c_node) pairs with r >= c in the spec's order: per column c, dz in {0,1},
    dy, dx in {-1,0,1}."""
    n = nx * ny * nz
    c = np.arange(n, dtype=np.int64)
    x = c % nx; y = (c // nx) % ny; z = c // (nx * ny)
    rows = []; cols = []; key = []
    slot = 0
    for dz in (0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = (x + dx >= 0) & (x + dx < nx) & (y + dy >= 0) & (y + dy < ny) & (z + dz < nz)
                r = c + dz * nx * ny + dy * nx + dx
                ok &= r >= c
                rows.append(r[ok]); cols.append(c[ok]); key.append(c[ok] * 32 + slot)
                slot += 1
    rows = np.concatenate(rows); cols = np.concatenate(cols); key = np.concatenate(key)
    order = np.argsort(key, kind="stable")
    return n, rows[order], cols[order]


def gen_stencil27_lower(nx: int, ny: int, nz: int):
    """3-D 27-point stencil, diag 26, off -1, symmetric lower, column-major."""
    n, r, c = _stencil27_pairs(nx, ny, nz)
    v = np.where(r == c, 26.0, -1.0)
    return n, r.astype(np.int32), c.astype(np.int32), v


def gen_elasticity_lower(nx: int, ny: int, nz: int):
    """3 dof/node, 27-pt node stencil x dense 3x3 blocks (SURVEY.md Appendix D generator):
    row = 3*node+dof; per column c=(node_c,dof_c): node pairs in stencil order, dof_r 0..2,
    keep r >= c; value 100 on the diagonal else -(1 + 0.25*((7r+13c) mod 4))."""
    nn, rn, cn = _stencil27_pairs(nx, ny, nz)
    n = 3 * nn
    # For column c = 3*cn+dc the file order is: node pairs (in stencil order) x dof_r.
    # Build per (pair, dc, dr) then order by (c, pair position within column, dr).
    npairs = rn.shape[0]
    pos = np.arange(npairs, dtype=np.int64)  # already sorted by (cn, slot)
    R = []; Cc = []; K = []
    for dc in range(3):
        for dr in range(3):
            r = 3 * rn + dr; c = 3 * cn + dc
            ok = r >= c
            R.append(r[ok]); Cc.append(c[ok]); K.append((c[ok] * npairs + pos[ok]) * 4 + dr)
    R = np.concatenate(R); Cc = np.concatenate(Cc); K = np.concatenate(K)
    order = np.argsort(K, kind="stable")
    R = R[order]; Cc = Cc[order]
    v = np.where(R == Cc, 100.0, -(1.0 + 0.25 * ((7 * R + 13 * Cc) % 4)))
    return n, R.astype(np.int32), Cc.astype(np.int32), v


def _mix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def gen_rmat(scale: int, edge_factor: int = 16, seed: int = 1, a=0.57, b=0.19, c=0.19,
             add_diagonal: bool = False):
    """R-MAT, general (unsymmetric).  Counter-based: edge e, level l draw u = hash(seed,e,l);
    duplicates summed; value U(-1,1) from hash(seed, e).  Returns row-sorted
    (then column-sorted) unique entries: n, I, J, V."""
    n = 1 << scale
    m = n * edge_factor
    e = np.arange(m, dtype=np.uint64)
    r = np.zeros(m, dtype=np.int64); cc = np.zeros(m, dtype=np.int64)
    with np.errstate(over="ignore"):
        for l in range(scale):
            h = _mix64(_mix64(e * np.uint64(64) + np.uint64(l)) ^ np.uint64(seed * 0x51ED27))
            u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
            qb = (u >= a) & (u < a + b)
            qc = (u >= a + b) & (u < a + b + c)
            qd = u >= a + b + c
            r = (r << 1) | (qc | qd)
            cc = (cc << 1) | (qb | qd)
        hv = _mix64(e ^ np.uint64((seed + 77) * 0x2545F491))
    v = (hv >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0
    if add_diagonal:
        d = np.arange(n, dtype=np.int64)
        r = np.concatenate([r, d]); cc = np.concatenate([cc, d]); v = np.concatenate([v, np.full(n, 4.0)])
    key = r * n + cc
    order = np.argsort(key, kind="stable")
    key = key[order]; v = v[order]
    uniq, start = np.unique(key, return_index=True)
    vs = np.add.reduceat(v, start)
    return n, (uniq // n).astype(np.int32), (uniq % n).astype(np.int32), vs


def write_mtx(path, n, i, j, v, symmetric: bool):
    """Matrix Market coordinate real {symmetric|general}, 1-based, the reader's fscanf format."""
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real %s\n" % ("symmetric" if symmetric else "general"))
        f.write("%d %d %d\n" % (n, n, len(i)))
        for a, b, c in zip(i.tolist(), j.tolist(), v.tolist()):
            f.write("%d %d %.17g\n" % (a + 1, b + 1, c))


# --------------------------------------------------------------------------------------
# the restatement
# --------------------------------------------------------------------------------------
class Oracle:
    def __init__(self):
        if not ORACLE_SO.exists():
            build()
        L = self.L = C.CDLL(str(ORACLE_SO))
        L.orc_fnv1a.restype = C.c_uint64
        L.orc_fnv1a.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_read_sym.restype = C.c_int
        L.orc_convert_fill.restype = C.c_longlong
        L.orc_csr_spmv_timed.restype = C.c_double
        L.orc_graph_unsym.restype = C.c_uint32
        L.orc_csr_spmv_timed.argtypes = [C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_int, C.c_int]

    def fnv(self, a: np.ndarray) -> str:
        a = np.ascontiguousarray(a)
        return "%016x" % self.L.orc_fnv1a(a.ctypes.data, a.nbytes)

    def num_threads(self) -> int:
        return int(self.L.orc_num_threads())

    def x_reference(self, n):
        x = np.empty(n, np.float64)
        self.L.orc_x_reference(C.c_int(n), _p(x, c_dbl_p))
        return x

    def read_sym(self, n, li, lj, lv, x=None):
        li = np.ascontiguousarray(li, np.int32); lj = np.ascontiguousarray(lj, np.int32)
        lv = np.ascontiguousarray(lv, np.float64)
        lower = len(li)
        tot = 2 * lower - n
        I = np.empty(tot, np.int32); J = np.empty(tot, np.int32); V = np.empty(tot, np.float64)
        rowIdx = np.empty(n + 1, np.int32); numInRow = np.empty(n, np.int32); diag = np.zeros(n, np.float64)
        maxCol = C.c_int(0)
        y = np.zeros(n, np.float64) if x is not None else None
        got = self.L.orc_read_sym(C.c_int(n), C.c_int(lower), _p(li, c_int_p), _p(lj, c_int_p), _p(lv, c_dbl_p),
                                  _p(I, c_int_p), _p(J, c_int_p), _p(V, c_dbl_p), _p(rowIdx, c_int_p),
                                  _p(numInRow, c_int_p), _p(diag, c_dbl_p), C.byref(maxCol),
                                  _p(x, c_dbl_p) if x is not None else None,
                                  _p(y, c_dbl_p) if x is not None else None)
        assert got == tot
        return dict(n=n, nnz=tot, I=I, J=J, V=V, rowIdx=rowIdx, numInRow=numInRow, diag=diag,
                    maxCol=maxCol.value, y=y, symmetric=True)

    def read_unsym(self, n, fi, fj, fv, x=None):
        fi = np.ascontiguousarray(fi, np.int32); fj = np.ascontiguousarray(fj, np.int32)
        fv = np.ascontiguousarray(fv, np.float64)
        tot = len(fi)
        I = np.empty(tot, np.int32); J = np.empty(tot, np.int32); V = np.empty(tot, np.float64)
        rowIdx = np.empty(n + 1, np.int32); numInRow = np.empty(n, np.int32)
        maxCol = C.c_int(0)
        y = np.zeros(n, np.float64) if x is not None else None
        self.L.orc_read_unsym(C.c_int(n), C.c_int(tot), _p(fi, c_int_p), _p(fj, c_int_p), _p(fv, c_dbl_p),
                              _p(I, c_int_p), _p(J, c_int_p), _p(V, c_dbl_p), _p(rowIdx, c_int_p),
                              _p(numInRow, c_int_p), C.byref(maxCol),
                              _p(x, c_dbl_p) if x is not None else None,
                              _p(y, c_dbl_p) if x is not None else None)
        return dict(n=n, nnz=tot, I=I, J=J, V=V, rowIdx=rowIdx, numInRow=numInRow, diag=np.zeros(n),
                    maxCol=maxCol.value, y=y, symmetric=False)

    def heuristic_ref(self, n, symmetric=True):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.L.orc_heuristic_ref(C.c_int(n), C.c_int(1 if symmetric else 0), C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def graph(self, m):
        n, tot = m["n"], m["nnz"]
        xadj = np.empty(n + 1, np.uint32)
        if m["symmetric"]:
            adj = np.empty(tot, np.uint32)
            self.L.orc_graph_sym(C.c_int(n), C.c_int(tot), _p(m["rowIdx"], c_int_p), _p(m["J"], c_int_p),
                                 _p(xadj, c_u32_p), _p(adj, c_u32_p))
            return xadj, adj
        adj = np.empty(2 * tot, np.uint32)
        k = self.L.orc_graph_unsym(C.c_int(n), C.c_int(tot), _p(m["I"], c_int_p), _p(m["J"], c_int_p),
                                   _p(xadj, c_u32_p), _p(adj, c_u32_p))
        return xadj, adj[:k].copy()

    def reorder(self, m, nParts, W, partVec):
        n, tot = m["n"], m["nnz"]
        partVec = np.ascontiguousarray(partVec, np.uint32)
        rowIdx = m["rowIdx"].copy(); numInRow = m["numInRow"].copy()
        numInRow2 = np.zeros(n, np.int32); pb = np.zeros(nParts + 1, np.int32); rl = np.zeros(n, np.int32)
        nI = np.empty(tot, np.int32); nJ = np.empty(tot, np.int32); nV = np.empty(tot, np.float64)
        self.L.orc_reorder(C.c_int(n), C.c_int(tot), C.c_int(nParts), C.c_int(W), _p(partVec, c_u32_p),
                           _p(m["I"], c_int_p), _p(m["J"], c_int_p), _p(m["V"], c_dbl_p), _p(rowIdx, c_int_p),
                           _p(numInRow, c_int_p), _p(numInRow2, c_int_p), _p(pb, c_int_p), _p(rl, c_int_p),
                           _p(nI, c_int_p), _p(nJ, c_int_p), _p(nV, c_dbl_p))
        return dict(n=n, nnz=tot, nParts=nParts, W=W, I=nI, J=nJ, V=nV, rowIdx=rowIdx, numInRow=numInRow,
                    numInRow2=numInRow2, partBoundary=pb, reorderList=rl, symmetric=m["symmetric"])

    def convert(self, r):
        n, P, W = r["n"], r["nParts"], r["W"]
        S = W // 32
        wE = np.zeros(P * S, np.int16); bE = np.zeros(P * S, np.int32)
        nER = np.zeros(n, np.int32); rlER = np.zeros(n, np.int32); rowVec = np.zeros(n, np.int32)
        wR = np.zeros(n // 32 + 2, np.int16); bR = np.zeros(n // 32 + 2, np.int32)
        realStart = np.zeros(P, np.int32); longRow = np.zeros(max(n, 1), np.int32)
        out = np.zeros(5, np.int64)
        self.L.orc_convert_plan(C.c_int(n), C.c_int(P), C.c_int(W), _p(r["partBoundary"], c_int_p),
                                _p(r["numInRow"], c_int_p), _p(r["numInRow2"], c_int_p),
                                _p(wE, c_i16_p), _p(bE, c_int_p), _p(nER, c_int_p), _p(rlER, c_int_p),
                                _p(rowVec, c_int_p), _p(wR, c_i16_p), _p(bR, c_int_p), _p(realStart, c_int_p),
                                _p(longRow, c_int_p), out.ctypes.data_as(C.POINTER(C.c_longlong)))
        sizeELL, sizeER, numOfRowER, toER, nLong = (int(v) for v in out)
        nb = (numOfRowER + 31) // 32
        colE = np.zeros(sizeELL, np.int16); valE = np.zeros(sizeELL, np.float64)
        colR = np.zeros(sizeER, np.int32); valR = np.zeros(sizeER, np.float64)
        waste = self.L.orc_convert_fill(C.c_int(n), C.c_int(P), C.c_int(W), _p(r["partBoundary"], c_int_p),
                                        _p(r["rowIdx"], c_int_p), _p(r["numInRow"], c_int_p),
                                        _p(r["I"], c_int_p), _p(r["J"], c_int_p), _p(r["V"], c_dbl_p),
                                        _p(wE, c_i16_p), _p(bE, c_int_p), _p(nER, c_int_p), _p(rlER, c_int_p),
                                        C.c_int(numOfRowER), _p(rowVec, c_int_p), _p(wR, c_i16_p), _p(bR, c_int_p),
                                        _p(realStart, c_int_p), _p(colE, c_i16_p), _p(valE, c_dbl_p),
                                        _p(colR, c_int_p), _p(valR, c_dbl_p))
        assert waste >= 0, "oracle converter hit a reference consistency abort"
        return dict(n=n, nParts=P, W=W, partBoundary=r["partBoundary"], widthVecBlockELL=wE, biasVecBlockELL=bE,
                    colBlockELL=colE, valBlockELL=valE, numOfRowER=numOfRowER, reorderListER=rlER,
                    rowVecER=rowVec[:numOfRowER].copy(), widthVecER=wR[:nb].copy(), biasVecER=bR[:nb].copy(),
                    colER=colR, valER=valR, sizeBlockELL=sizeELL, sizeER=sizeER, toER=toER, wasteElement=int(waste),
                    nLongVec=nLong, longRow=longRow[:nLong].copy(), numInRowER=nER)

    def emulate(self, e, r, x, use_fma=True):
        n = e["n"]
        y = np.zeros(n, np.float64)
        x = np.ascontiguousarray(x, np.float64)
        self.L.orc_emulate(C.c_int(n), C.c_int(e["nParts"]), C.c_int(e["W"]), _p(e["partBoundary"], c_int_p),
                           _p(e["widthVecBlockELL"], c_i16_p), _p(e["biasVecBlockELL"], c_int_p),
                           _p(e["colBlockELL"], c_i16_p), _p(e["valBlockELL"], c_dbl_p),
                           C.c_int(e["numOfRowER"]), _p(e["rowVecER"], c_int_p), _p(e["widthVecER"], c_i16_p),
                           _p(e["biasVecER"], c_int_p), _p(e["colER"], c_int_p), _p(e["valER"], c_dbl_p),
                           C.c_int(e["nLongVec"]), _p(e["longRow"], c_int_p), _p(r["rowIdx"], c_int_p),
                           _p(r["J"], c_int_p), _p(r["V"], c_dbl_p), _p(x, c_dbl_p), _p(y, c_dbl_p),
                           C.c_int(1 if use_fma else 0))
        return y

    def csr_spmv(self, rowIdx, J, V, x):
        n = len(rowIdx) - 1
        y = np.empty(n, np.float64)
        self.L.orc_csr_spmv(C.c_int(n), _p(rowIdx, c_int_p), _p(J, c_int_p), _p(V, c_dbl_p),
                            _p(np.ascontiguousarray(x, np.float64), c_dbl_p), _p(y, c_dbl_p))
        return y

    def csr_abs_spmv(self, rowIdx, J, V, x):
        n = len(rowIdx) - 1
        y = np.empty(n, np.float64)
        self.L.orc_csr_abs_spmv(C.c_int(n), _p(rowIdx, c_int_p), _p(J, c_int_p), _p(V, c_dbl_p),
                                _p(np.ascontiguousarray(x, np.float64), c_dbl_p), _p(y, c_dbl_p))
        return y

    def csr_spmv_timed(self, rowIdx, J, V, x, warmup, iters):
        n = len(rowIdx) - 1
        y = np.empty(n, np.float64)
        x = np.ascontiguousarray(x, np.float64)
        return float(self.L.orc_csr_spmv_timed(n, _p(rowIdx, c_int_p), _p(J, c_int_p), _p(V, c_dbl_p),
                                               _p(x, c_dbl_p), _p(y, c_dbl_p), warmup, iters)), y

    def x_of_global(self, ids, scale=1.0, shift=0.0):
        """x_j = hash(j) * scale + shift for natural indices j (the function multigpu.x_of_global)."""
        ids = np.ascontiguousarray(ids, np.int64)
        x = np.empty(len(ids))
        self.L.orc_x_of_global_fill(C.c_longlong(len(ids)), _p(ids, C.POINTER(C.c_longlong)), C.c_double(scale), C.c_double(shift), _p(x, c_dbl_p))
        return x

    def stencil27_rows_product(self, grid, ids, scale=1.0, shift=0.0):
        """(y_ref, |A||x|) of the rows `ids` (natural indices) of the 27-point stencil on `grid`, in
        closed form, for x_j = hash(j) * scale + shift: the check vector of BASELINE.json config 5."""
        ids = np.ascontiguousarray(ids, np.int64)
        y = np.empty(len(ids)); a = np.empty(len(ids))
        self.L.orc_stencil27_rows_product(int(grid[0]), int(grid[1]), int(grid[2]), C.c_longlong(len(ids)), _p(ids, C.POINTER(C.c_longlong)),
                                          C.c_double(scale), C.c_double(shift), _p(y, c_dbl_p), _p(a, c_dbl_p))
        return y, a

    def vector_reorder(self, v, lst):
        out = np.empty_like(v)
        self.L.orc_vector_reorder(C.c_int(len(v)), _p(v, c_dbl_p), _p(out, c_dbl_p), _p(lst, c_int_p))
        return out

    def vector_recover(self, vr, lst):
        out = np.empty_like(vr)
        self.L.orc_vector_recover(C.c_int(len(vr)), _p(vr, c_dbl_p), _p(out, c_dbl_p), _p(lst, c_int_p))
        return out


def csr_from_sorted_coo(m):
    """Row-sorted COO (symmetric reader / any reordered matrix) is already CSR."""
    return m["rowIdx"], m["J"], m["V"]


# --------------------------------------------------------------------------------------
# the unmodified reference
# --------------------------------------------------------------------------------------
def mtmetis_partition(xadj, adj, nparts, nthreads=1, ub=1.001):
    """Run the pinned mt-metis binary through bin/ehyb_mtmetis (same call as reordering.c:280)."""
    import tempfile
    n = len(xadj) - 1
    with tempfile.TemporaryDirectory() as d:
        g = os.path.join(d, "g.bin"); w = os.path.join(d, "w.bin")
        with open(g, "wb") as f:
            f.write(np.array([0x47594845, n, nparts, nthreads], np.uint32).tobytes())
            f.write(np.array([ub], np.float32).tobytes())
            f.write(np.ascontiguousarray(xadj, np.uint32).tobytes())
            f.write(np.ascontiguousarray(adj, np.uint32).tobytes())
        subprocess.check_call([str(MTMETIS_BIN), g, w])
        raw = np.fromfile(w, dtype=np.uint32)
    return raw[1:].copy()


class Reference:
    """oracle/_ref/libehyb_ref.so: reordering.c, convert.c, mmio.c and solver_test.c's readers,
    compiled unmodified."""

    def __init__(self):
        if not REF_SO.exists():
            raise FileNotFoundError(str(REF_SO))
        L = self.L = C.CDLL(str(REF_SO))
        L.ref_sizeof_matrixCOO.restype = C.c_size_t
        L.ref_sizeof_matrixEHYB.restype = C.c_size_t
        L.ref_last_partition.restype = C.c_uint32
        assert L.ref_sizeof_matrixCOO() == C.sizeof(MatrixCOO)
        assert L.ref_sizeof_matrixEHYB() == C.sizeof(MatrixEHYB)
        if MTMETIS_BIN.exists():
            L.ref_set_mtmetis_bin(str(MTMETIS_BIN).encode())
        self._keep = None

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def read_mtx(self, path):
        coo = MatrixCOO(); x = c_dbl_p(); y = c_dbl_p(); sym = C.c_int()
        rc = self.L.ref_read_mtx(str(path).encode(), C.byref(coo), C.byref(x), C.byref(y), C.byref(sym))
        if rc:
            raise RuntimeError("reference reader failed: %d" % rc)
        n, tot = coo.dimension, coo.totalNum
        m = dict(n=n, nnz=tot, I=np_from(coo.I, tot, np.int32), J=np_from(coo.J, tot, np.int32),
                 V=np_from(coo.V, tot, np.float64), rowIdx=np_from(coo.rowIdx, n + 1, np.int32),
                 numInRow=np_from(coo.numInRow, n, np.int32), maxCol=coo.maxCol, nParts=coo.nParts,
                 W=coo.vectorCacheSize, kpp=coo.kernelPerPart, x=np_from(x, n, np.float64),
                 y=np_from(y, n, np.float64), symmetric=bool(sym.value))
        return m

    def pipeline(self, m, nParts, W, kpp=1, partVec=None):
        """reader state -> matrixReorder[_unsym] -> COO2EHYB, all reference code.
        partVec=None runs the pinned mt-metis binary exactly as the reference calls it."""
        n, tot = m["n"], m["nnz"]
        coo = MatrixCOO()
        self.L.ref_coo_from_arrays(C.byref(coo), n, tot, _p(m["I"], c_int_p), _p(m["J"], c_int_p),
                                   _p(m["V"], c_dbl_p), _p(m["rowIdx"], c_int_p), _p(m["numInRow"], c_int_p),
                                   m["maxCol"], nParts, W, kpp)
        if partVec is not None:
            pv = np.ascontiguousarray(partVec, np.uint32)
            self._keep = pv
            self.L.ref_set_partition(_p(pv, c_u32_p), C.c_uint32(n))
        else:
            self.L.ref_set_partition(None, C.c_uint32(0))
        rc = self.L.ref_matrixReorder(C.byref(coo), 1 if m["symmetric"] else 0)
        if rc:
            raise RuntimeError("reference matrixReorder exited with %d" % (rc - 1000))
        where = np.zeros(n, np.uint32)
        self.L.ref_last_partition(_p(where, c_u32_p), C.c_uint32(n))
        info = np.zeros(8, np.uint64)
        self.L.ref_graph_info(info.ctypes.data_as(C.POINTER(C.c_uint64)))
        r = dict(n=n, nnz=tot, nParts=nParts, W=W, I=np_from(coo.I, tot, np.int32), J=np_from(coo.J, tot, np.int32),
                 V=np_from(coo.V, tot, np.float64), rowIdx=np_from(coo.rowIdx, n + 1, np.int32),
                 numInRow=np_from(coo.numInRow, n, np.int32), numInRow2=np_from(coo.numInRow2, n, np.int32),
                 partBoundary=np_from(coo.partBoundary, nParts + 1, np.int32),
                 reorderList=np_from(coo.reorderList, n, np.int32), partVec=where, graph_info=info,
                 symmetric=m["symmetric"])
        eh = MatrixEHYB(); sE = C.c_int(); sR = C.c_int()
        rc = self.L.ref_COO2EHYB(C.byref(coo), C.byref(eh), C.byref(sE), C.byref(sR))
        if rc:
            r["convert_exit"] = rc - 1000
            return r, None
        S = W // 32
        nER = eh.numOfRowER
        nb = (nER + 31) // 32
        e = dict(n=n, nParts=nParts, W=W, partBoundary=r["partBoundary"],
                 widthVecBlockELL=np_from(eh.widthVecBlockELL, nParts * S, np.int16),
                 biasVecBlockELL=np_from(eh.biasVecBlockELL, nParts * S, np.int32),
                 colBlockELL=np_from(eh.colBlockELL, sE.value, np.int16),
                 valBlockELL=np_from(eh.valBlockELL, sE.value, np.float64),
                 numOfRowER=nER, reorderListER=np_from(eh.reorderListER, n, np.int32),
                 rowVecER=np_from(eh.rowVecER, nER, np.int32), widthVecER=np_from(eh.widthVecER, nb, np.int16),
                 biasVecER=np_from(eh.biasVecER, nb, np.int32), colER=np_from(eh.colER, sR.value, np.int32),
                 valER=np_from(eh.valER, sR.value, np.float64), sizeBlockELL=sE.value, sizeER=sR.value,
                 nLongVec=eh.nLongVec, longRow=np.zeros(0, np.int32))
        return r, e
