"""Host side of the overflow stream (csrc/host/ovfstream.c): the tile-packed device format, decoded here
by a lane-by-lane numpy restatement of ehyb_ovfstream_kernel + ehyb_ovfstream_fixup (same segment
arithmetic from the group masks, same carry slots, same "who stores what"), must give the row sums of
the list.  Runs without a GPU: it pins the FORMAT and the kernel's bookkeeping; the GPU tier
(test_gpu_parity.py::test_overflow_stream_is_deterministic, test_gpu_configs.py config 4) pins the kernel.
"""
import ctypes as C

import numpy as np
import pytest

from ehyb_spmv_gpu_b200 import _lib as L

HUB = 0x80000000


class OvfStream(C.Structure):
    _fields_ = [("count", C.c_int64), ("nTiles", C.c_int64), ("nSeg", C.c_int64), ("hubRefs", C.c_int64), ("deviceBytes", C.c_int64),
                ("nHub", C.c_int), ("tileGroups", C.c_int), ("tileBytes", C.c_int),
                ("tiles", C.POINTER(C.c_ubyte)), ("rowOfSeg", C.POINTER(C.c_int32)), ("hubCols", C.POINTER(C.c_int32)),
                ("carryRow", C.POINTER(C.c_int32)), ("nRuns", C.c_int64), ("nRunsShort", C.c_int64), ("runs", C.POINTER(C.c_int32))]


def _build(row, col, val, ncols, hub_cap, tg):
    lib = L.load()
    st = OvfStream()
    f = lib.ehyb_ovfstream_build
    f.restype = C.c_int
    f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(OvfStream)]
    rc = f(len(row), row.ctypes.data, col.ctypes.data, val.ctypes.data, ncols, hub_cap, tg, C.byref(st))
    return rc, st


def _emulate(st, x, n, accumulate_into=None):
    """What the kernels do, lane by lane (sums in the kernel's order: Hillis-Steele inside a group)."""
    TG, E, tb = st.tileGroups, 32 * st.tileGroups, st.tileBytes
    assert tb == 12 * E + 8 * TG + 16
    blob = np.ctypeslib.as_array(st.tiles, shape=(st.nTiles * tb,))
    rowOfSeg = np.ctypeslib.as_array(st.rowOfSeg, shape=(st.nSeg,))
    hubCols = np.ctypeslib.as_array(st.hubCols, shape=(max(st.nHub, 1),))[:st.nHub] if st.nHub else np.zeros(0, np.int32)
    carryRow = np.ctypeslib.as_array(st.carryRow, shape=(2 * st.nTiles,))
    hub = x[hubCols] if st.nHub else np.zeros(0)
    y = np.zeros(n) if accumulate_into is None else accumulate_into.copy()
    acc = accumulate_into is not None
    written = np.zeros(n, bool)
    carryVal = np.zeros(2 * st.nTiles)
    for t in range(st.nTiles):
        rec = blob[t * tb:(t + 1) * tb]
        v = rec[:8 * E].view(np.float64)
        c = rec[8 * E:12 * E].view(np.uint32)
        g = rec[12 * E:12 * E + 8 * TG].view(np.uint32).reshape(TG, 2)
        flags = int(rec[12 * E + 8 * TG:12 * E + 8 * TG + 4].view(np.uint32)[0])
        headCont, tailCont = bool(flags & 1), bool(flags & 2)
        assert TG == 4, "the kernel's lane mapping: four consecutive entries per lane"
        headSeg = int(g[0, 0])
        tailSeg = int(g[3, 0]) + bin(int(g[3, 1]) & ~1).count("1")

        def store(sg, sm):
            r = int(rowOfSeg[sg])
            if r < 0:
                return
            isHead, isTail = headCont and sg == headSeg, tailCont and sg == tailSeg
            if isHead:
                assert carryRow[2 * t] == r
                if isTail:
                    assert carryRow[2 * t + 1] == r
                carryVal[2 * t] = sm
            elif isTail:
                assert carryRow[2 * t + 1] == r
                carryVal[2 * t + 1] = sm
            else:
                assert not written[r], "two stores to one row"
                written[r] = True
                y[r] = y[r] + sm if acc else sm

        # lane j owns entries 4j..4j+3: serial sums inside the lane, one segmented scan across the lanes
        run = np.zeros(32); head = np.zeros(32); nStart = np.zeros(32, int); sB = np.zeros(32, int); nibs = np.zeros(32, int)
        for lane in range(32):
            gmx, gmask = int(g[lane >> 3, 0]), int(g[lane >> 3, 1])
            sh = 4 * (lane & 7)
            nib = (gmask >> sh) & 0xF
            seg0 = gmx + bin(gmask & ((2 << sh) - 1) & ~1).count("1")
            sB[lane] = seg0 - (nib & 1); nibs[lane] = nib; nStart[lane] = bin(nib).count("1")
            r_, h_, seen = 0.0, 0.0, 0
            for k in range(4):
                cc = int(c[4 * lane + k])
                xv = hub[cc & 0x7FFFFFFF] if cc & HUB else x[cc]
                pk = xv * v[4 * lane + k]
                if (nib >> k) & 1:
                    if seen == 0:
                        h_ = r_
                    else:
                        store(sB[lane] + seen, r_)
                    r_ = 0.0
                    seen += 1
                r_ += pk
            run[lane], head[lane] = r_, h_
        starts = [l for l in range(32) if nStart[l] > 0]
        dist = np.array([lane - max([l for l in starts if l <= lane], default=0) for lane in range(32)])
        sc = run.copy()
        off = 1
        while off < 32:
            shv = np.concatenate([np.zeros(off), sc[:-off]])
            sc = np.where(dist >= off, sc + shv, sc)
            off <<= 1
        before = np.concatenate([[0.0], sc[:-1]])
        for lane in range(32):
            if nStart[lane] > 0 and not (lane == 0 and (nibs[lane] & 1)):
                store(sB[lane], before[lane] + head[lane])
        store(tailSeg, sc[31])
    # fix-up (ehyb_ovfstream_fixup): the runs the builder lists - short ones by one thread, long ones by a warp
    # (lane l adds slots l, l + 32, ... in order, then a fixed xor-shuffle tree)
    runs = np.ctypeslib.as_array(st.runs, shape=(max(st.nRuns, 1), 3))[:st.nRuns]
    seen = np.zeros(2 * st.nTiles, bool)
    for k, (first, ln, r) in enumerate(runs):
        assert ln >= 1 and np.all(carryRow[first:first + ln] == r) and r >= 0
        assert (first == 0 or carryRow[first - 1] != r) and (first + ln == 2 * st.nTiles or carryRow[first + ln] != r), "not a maximal run"
        assert (ln <= 16) == (k < st.nRunsShort), "short runs first"
        assert not seen[first:first + ln].any()
        seen[first:first + ln] = True
        if k < st.nRunsShort:
            sm = 0.0
            for q in range(ln):
                sm += carryVal[first + q]
        else:
            lanes = np.zeros(32)
            for q in range(ln):
                lanes[q % 32] += carryVal[first + q]
            off = 16
            while off:
                lanes = lanes + lanes[np.arange(32) ^ off]
                off >>= 1
            sm = lanes[0]
        assert not written[r], "a row is both stored and fixed up"
        written[r] = True
        y[r] = y[r] + sm if acc else sm
    assert np.array_equal(seen, carryRow >= 0), "every used carry slot is in exactly one run"
    return y, written


def _list(kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "powerlaw":
        n = 3000
        deg = np.minimum((rng.pareto(1.2, n) * 4).astype(np.int64), 900)
        deg[rng.integers(0, n, n // 3)] = 0               # empty rows
        row = np.repeat(np.arange(n, dtype=np.int32), deg)
        col = np.minimum((rng.pareto(1.0, len(row)) * 20).astype(np.int64), n - 1).astype(np.int32)
    elif kind == "long_row":
        n = 2000
        row = np.concatenate([np.zeros(1500, np.int32), np.arange(1, n, dtype=np.int32), np.full(700, n - 1, np.int32)])
        row.sort()
        col = rng.integers(0, n, len(row)).astype(np.int32)
    elif kind == "exact_tiles":                                   # a multiple of the tile size: no padding segment
        n = 512
        row = np.repeat(np.arange(n, dtype=np.int32), 4)          # 2 048 entries, rows end on group and tile borders
        col = rng.integers(0, n, len(row)).astype(np.int32)
    else:                                                         # one entry
        n = 10
        row = np.array([7], np.int32)
        col = np.array([3], np.int32)
    val = rng.uniform(-1, 1, len(row))
    return n, row, col, val


@pytest.mark.parametrize("kind", ["powerlaw", "long_row", "exact_tiles", "single"])
@pytest.mark.parametrize("tg", [4])
@pytest.mark.parametrize("hub_cap", [0, 64])
def test_tile_stream_decodes_to_the_row_sums(kind, tg, hub_cap):
    n, row, col, val = _list(kind, 3)
    rc, st = _build(row, col, val, n, hub_cap, tg)
    assert rc == 0
    try:
        E = 32 * tg
        assert st.count == len(row) and st.nTiles == (len(row) + E - 1) // E and st.tileBytes % 16 == 0
        assert st.nSeg == len(np.unique(row)) + (1 if len(row) % E else 0)
        if hub_cap:
            hubs = np.ctypeslib.as_array(st.hubCols, shape=(st.nHub,)) if st.nHub else np.zeros(0, np.int32)
            cnt = np.bincount(col, minlength=n)
            assert st.nHub <= hub_cap and np.all(cnt[hubs] >= 2)
            if st.nHub and st.nHub < np.count_nonzero(cnt >= 2):    # no column left out is referenced more than a hub
                rest = np.setdiff1d(np.arange(n), hubs)
                assert cnt[rest].max() <= cnt[hubs].min()
            assert st.hubRefs == int(cnt[hubs].sum())
        else:
            assert st.nHub == 0 and st.hubRefs == 0
        x = np.random.default_rng(9).uniform(-1, 1, n)
        y, written = _emulate(st, x, n)
        ref = np.zeros(n)
        np.add.at(ref, row, val * x[col])
        assert np.array_equal(written, np.bincount(row, minlength=n) > 0), "exactly the non-empty rows are written, once"
        assert np.allclose(y, ref, rtol=0, atol=1e-12 * max(1.0, np.abs(ref).max()))
        base = np.random.default_rng(4).uniform(-1, 1, n)            # accumulate mode (the main kernel wrote y first)
        y2, _ = _emulate(st, x, n, accumulate_into=base)
        assert np.allclose(y2, base + ref, rtol=0, atol=1e-12 * max(1.0, np.abs(ref).max()))
    finally:
        L.load().ehyb_ovfstream_free(C.byref(st))


@pytest.mark.parametrize("seed", range(24))
def test_tile_stream_random_lists(seed):
    """Randomized shapes: row lengths from 0 to several tiles, row starts on lane, group and tile borders,
    list lengths around multiples of the tile size, random hub capacities - decoded by the kernel's
    bookkeeping, every non-empty row written exactly once with its sum."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(5, 400))
    kind = seed % 4
    if kind == 0:
        deg = rng.integers(0, 9, n)                                   # short rows: many starts per lane
    elif kind == 1:
        deg = np.where(rng.random(n) < 0.1, rng.integers(100, 700, n), rng.integers(0, 3, n))   # a few rows over several tiles
    elif kind == 2:
        deg = np.full(n, 4 * int(rng.integers(1, 9)))                 # starts aligned with the lanes' four entries
    else:
        deg = rng.integers(0, 70, n)
    deg[rng.integers(0, n)] += 1                                      # at least one entry
    total = int(deg.sum())
    if seed % 3 == 0:                                                 # land exactly on a tile border
        deg[-1] += (-total) % 128
    row = np.repeat(np.arange(n, dtype=np.int32), deg)
    col = rng.integers(0, n, len(row)).astype(np.int32)
    val = rng.uniform(-1, 1, len(row))
    hub_cap = int(rng.choice([0, 1, 7, 64, 1000]))
    rc, st = _build(row, col, val, n, hub_cap, 4)
    assert rc == 0
    try:
        x = rng.uniform(-1, 1, n)
        y, written = _emulate(st, x, n)
        ref = np.zeros(n)
        np.add.at(ref, row, val * x[col])
        assert np.array_equal(written, deg > 0)
        assert np.allclose(y, ref, rtol=0, atol=1e-12 * max(1.0, float(np.abs(ref).max())))
    finally:
        L.load().ehyb_ovfstream_free(C.byref(st))


def test_column_blocks_partition_the_list():
    """ehyb_ovfstream_build_blocked: every entry in exactly one block (by column range), the blocks decoded
    one after the other and added up give the row sums; an empty block has no arrays."""
    n, row, col, val = _list("powerlaw", 5)
    col = np.where(col >= 2 * (n // 4), np.minimum(col + n // 4, n - 1), col).astype(np.int32)   # nothing in block 2 of 4
    lib = L.load()
    nb, bc = 4, (n + 3) // 4
    sts = (OvfStream * nb)()
    f = lib.ehyb_ovfstream_build_blocked
    f.restype = C.c_int
    f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(OvfStream)]
    assert f(len(row), row.ctypes.data, col.ctypes.data, val.ctypes.data, n, 16, 4, nb, bc, sts) == 0
    try:
        assert sum(st.count for st in sts) == len(row)
        x = np.random.default_rng(9).uniform(-1, 1, n)
        y = np.zeros(n)
        for b, st in enumerate(sts):
            inb = (col // bc) == b
            assert st.count == int(inb.sum())
            if st.count == 0:
                assert not st.tiles
                continue
            y, _ = _emulate(st, x, n, accumulate_into=y)
        ref = np.zeros(n)
        np.add.at(ref, row, val * x[col])
        assert sts[2].count == 0 and sts[0].count > 0
        assert np.allclose(y, ref, rtol=0, atol=1e-12 * max(1.0, np.abs(ref).max()))
    finally:
        for st in sts:
            lib.ehyb_ovfstream_free(C.byref(st))


def test_unsorted_list_is_refused():
    row = np.array([0, 2, 1], np.int32); col = np.zeros(3, np.int32); val = np.ones(3)
    rc, st = _build(row, col, val, 4, 0, 4)
    assert rc != 0
