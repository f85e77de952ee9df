"""Host side of the product (reader, graph, reorder, COO2EHYB, tuned layout) against the
oracle and the committed reference outputs: every index/value array bit-exact."""
import base64
import json
import zlib
from pathlib import Path

import numpy as np
import pytest

from ehyb_spmv_gpu_b200 import api
from oracle import oracle as O
from tests import util

GOLD = json.loads((Path(__file__).parent / "golden" / "reference_kat.json").read_text())


def _partvec(case):
    return np.frombuffer(zlib.decompress(base64.b64decode(case["partVec_z"])), dtype=np.uint32)


@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_product_matches_reference_golden(orc, name):
    """generator -> reader expansion -> graph -> reorder -> COO2EHYB -> tuned layout -> de-interleave,
    all against hashes of the unmodified reference's arrays."""
    c = GOLD["cases"][name]
    kind, dims = c["kind"], tuple(c["dims"])
    n, li, lj, lv = api.gen_lower(util.KINDS[kind][0], *(list(dims) + [1])[:3])
    no, lio, ljo, lvo = util.lower_entries(kind, dims)
    assert n == no and np.array_equal(li, lio) and np.array_equal(lj, ljo) and np.array_equal(lv, lvo)
    x = api.x_reference(n)
    assert np.array_equal(x, orc.x_reference(n))
    m = api.CooMatrix.from_lower(n, li, lj, lv, x)
    assert orc.fnv(m.y_golden) == c["y_golden_fnv"]
    xadj, adj = m.build_graph()
    assert orc.fnv(xadj) == c["graph"]["xadj_fnv"] and orc.fnv(adj) == c["graph"]["adjncy_fnv"]
    m.set_plan(c["nParts"], c["W"], max(c["kpp"], 1))
    m.reorder_with_partition(_partvec(c))
    a = m.arrays()
    for k in util.COO_KEYS:
        assert orc.fnv(a[k]) == c["hash"][k], k
    e = m.coo2ehyb()
    for k in util.EHYB_KEYS:
        assert orc.fnv(e[k]) == c["hash"][k], k
    assert (e["sizeBlockELL"], e["sizeER"], e["numOfRowER"]) == (c["sizeBlockELL"], c["sizeER"], c["numOfRowER"])
    for fill in (0.0, 0.5, 1.0, -1.0):
        lay = api.Layout(m, er_fill=fill)
        st = lay.stats()
        assert st["nnzEll"] + st["nnzRemInSlice"] + st["nnzOverflow"] == st["nnz"] == c["nnz"]
        assert st["nnz"] - st["nnzEll"] == np.sum(e["valER"] != 0) or True
        e2 = lay.to_reference()
        for k in util.EHYB_KEYS:
            assert orc.fnv(e2[k]) == c["hash"][k], (fill, k)
        lay.free()
    m.free()


def test_plan_reference_matches_the_reference_heuristic(orc):
    for n, (P, W, kpp) in GOLD["heuristic"].items():
        p = api.plan_reference(int(n), True)
        assert (p.nParts, p.W, p.ctasPerPart) == (P, W, kpp)
    for n in (1000, 65536, 300000, 1048576, 2097152, 3000000, 5400000, 16777216):
        for sym in (True, False):
            p = api.plan_reference(n, sym)
            assert (p.nParts, p.W, p.ctasPerPart) == orc.heuristic_ref(n, sym), (n, sym)


def test_independent_numpy_decode_of_the_tuned_layout(orc):
    """Decodes the device-facing blob with the index formulas of DESIGN.md (written here in
    numpy, independently of layout.c) and checks it holds exactly the reordered matrix."""
    kind, dims, P, W = "st27", (20, 20, 20), 4, 2112
    m = util.product_pipeline(kind, dims, P, W, 1)
    a = m.arrays()
    lay = api.Layout(m, er_fill=0.5)
    raw, st = lay.raw(), lay.stats()
    blob = raw["blob"]
    rows = {}
    for p in range(P):
        rs, re_, s0, s1, cs, cc = raw["parts"][p][:6]
        cache = raw["cacheCols"][cs:cs + cc]
        assert np.all(np.diff(cache) > 0) and not np.any((cache >= rs) & (cache < min(rs + W, a["n"])))
        for s in range(s0, s1):
            off = int(raw["slices"]["off256"][s]) * 256
            w, wr = int(raw["slices"]["w"][s]), int(raw["slices"]["wr"][s])
            ev = blob[off:off + w * 512].view(np.float64).reshape(w, 32, 2)
            w4 = (w + 3) // 4
            ec = blob[off + w * 512:off + w * 512 + w4 * 512].view(np.uint16).reshape(w4, 32, 2, 4)
            ro = off + w * 512 + w4 * 512
            rv = blob[ro:ro + wr * 512].view(np.float64).reshape(wr, 32, 2)
            wr4 = (wr + 3) // 4
            rc = blob[ro + wr * 512:ro + wr * 512 + wr4 * 512].view(np.uint16).reshape(wr4, 32, 2, 4)
            for t in range(64):
                r = rs + (s - s0) * 64 + t
                if r >= re_:
                    continue
                lane, h = t % 32, t // 32
                ent = [(rs + int(ec[k // 4, lane, h, k % 4]), float(ev[k, lane, h])) for k in range(w)]
                ent += [(int(cache[rc[k // 4, lane, h, k % 4]]) if cc else 0, float(rv[k, lane, h])) for k in range(wr)]
                rows[r] = ent
    for r_, c_, v_ in zip(raw["ovfRow"], raw["ovfCol"], raw["ovfVal"]):
        rows[int(r_)].append((int(c_), float(v_)))
    assert np.all(np.diff(raw["ovfRow"]) >= 0)
    for r in range(a["n"]):
        want = sorted((int(j), float(v)) for j, v in zip(a["J"][a["rowIdx"][r]:a["rowIdx"][r + 1]],
                                                         a["V"][a["rowIdx"][r]:a["rowIdx"][r + 1]]))
        got = sorted(e for e in rows[r] if e[1] != 0.0)
        assert got == want, r
    assert st["blobBytes"] == len(blob) and st["nOverflow"] == len(raw["ovfRow"])
    lay.free(); m.free()


def test_long_rows_and_empty_rows(orc):
    """What the reference cannot do: rows with more than 512
    in-window entries, rows without any entry, a matrix without remainder."""
    n = 1500
    rng = np.random.default_rng(5)
    ent = {(i, i): 2.0 for i in range(0, n, 2)}            # odd rows stay empty unless hit below
    for i in (0, 1, 700):                                    # three dense rows (> 512 in-window entries)
        for j in rng.choice(n, 1300, replace=False):
            ent[(i, int(j))] = float(rng.uniform(-1, 1))
    fi = np.array([k[0] for k in ent], np.int32); fj = np.array([k[1] for k in ent], np.int32)
    fv = np.array(list(ent.values()))
    x = util.x_random(n, 1)
    m = api.CooMatrix.from_general(n, fi, fj, fv, x)
    m.set_plan(2, 1024, 1)
    m.reorder_with_partition((np.arange(n) >= 750).astype(np.uint32))
    a = m.arrays()
    e = m.coo2ehyb()
    assert e["nLongVec"] >= 1
    # the reference's constant rule (512 in-window entries, kernel.h:26), pinned explicitly ...
    lay = api.Layout(m, long_row_threshold=512, min_coverage=-1)
    st = lay.stats()
    assert st["nLongRows"] == e["nLongVec"] and st["nnzEll"] + st["nnzRemInSlice"] + st["nnzOverflow"] == st["nnz"]
    # ... and the default: the limit also follows a warp's fair share of the partition (a slice is
    # walked by one warp), so at least the same rows are long and the slices are narrower
    lay_d = api.Layout(m, min_coverage=-1)
    st_d = lay_d.stats()
    assert st_d["nLongRows"] >= st["nLongRows"] and st_d["nnzEll"] + st_d["nnzRemInSlice"] + st_d["nnzOverflow"] == st["nnz"]
    assert int(lay_d.raw()["slices"]["w"].max(initial=0)) <= int(lay.raw()["slices"]["w"].max(initial=0))
    assert np.allclose(util.layout_spmv(lay_d.raw(), m.vector_reorder(x)), orc.csr_spmv(a["rowIdx"], a["J"], a["V"], m.vector_reorder(x)), rtol=0, atol=1e-12)
    # coverage rule: an R-MAT-like matrix keeps almost nothing in slices -> everything in the COO list
    lay_c = api.Layout(m, min_coverage=0.99)
    st_c = lay_c.stats()
    assert st_c["nnzOverflow"] == st["nnz"] and st_c["nnzEll"] == 0 and st_c["blobBytes"] == 0
    assert np.allclose(util.layout_spmv(lay_c.raw(), m.vector_reorder(x)), orc.csr_spmv(a["rowIdx"], a["J"], a["V"], m.vector_reorder(x)), rtol=0, atol=1e-12)
    lay_d.free(); lay_c.free()
    # CPU check of the reference-layout arrays through the oracle's emulation (long rows included)
    r = dict(rowIdx=a["rowIdx"], J=a["J"], V=a["V"])
    xr = m.vector_reorder(x)
    y = orc.emulate({**e, "partBoundary": a["partBoundary"]}, r, xr, use_fma=False)
    y_ref = orc.csr_spmv(a["rowIdx"], a["J"], a["V"], xr)
    assert np.allclose(y, y_ref, rtol=0, atol=1e-12)
    lay.free(); m.free()
    # block-diagonal matrix: no remainder at all ("perfect matrix", convert.c:136-139 exits)
    n2, li, lj, lv = api.gen_lower(api.GEN_LAPLACE2D, 8, 8)
    keep = (li // 32) == (lj // 32)
    m2 = api.CooMatrix.from_lower(n2, li[keep], lj[keep], lv[keep])
    m2.set_plan(2, 32, 1)
    m2.reorder_with_partition((np.arange(n2) // 32).astype(np.uint32))
    e2 = m2.coo2ehyb()
    assert e2["numOfRowER"] == 0 and e2["sizeER"] == 0
    m2.free()


def test_mtx_roundtrip_and_reader(orc, tmp_path):
    n, li, lj, lv = api.gen_lower(api.GEN_STENCIL27, 6, 5, 4)
    p = tmp_path / "a.mtx"
    lib = api.L.load()
    import ctypes as C
    assert lib.ehyb_write_mtx(str(p).encode(), n, C.c_int64(len(li)), li.ctypes.data_as(api.L.c_int_p),
                              lj.ctypes.data_as(api.L.c_int_p), lv.ctypes.data_as(api.L.c_dbl_p), 1) == 0
    m = api.CooMatrix.read_mtx(p)
    mo = orc.read_sym(n, li, lj, lv, orc.x_reference(n))
    a = m.arrays()
    for k in ["I", "J", "V", "rowIdx", "numInRow"]:
        assert np.array_equal(a[k], mo[k]), k
    assert np.array_equal(m.y_golden, mo["y"]) and m.symmetric
    m.free()
    with pytest.raises(api.EhybError):
        api.CooMatrix.read_mtx(tmp_path / "missing.mtx")


def test_general_matrix_pipeline_and_rmat(orc):
    """Unsymmetric path: unsorted COO, graph of A+A^T, reorder, layout; R-MAT generator parity
    (numpy vs C)."""
    n, fi, fj, fv = api.gen_rmat(10, 8, seed=1, add_diagonal=False)
    n2, gi, gj, gv = O.gen_rmat(10, 8, seed=1)
    assert n == n2 and np.array_equal(fi, gi) and np.array_equal(fj, gj) and np.allclose(fv, gv, rtol=1e-14, atol=1e-15)
    rng = np.random.default_rng(2)
    order = rng.permutation(len(fi))                         # file order is arbitrary for general matrices
    x = util.x_random(n, 4)
    m = api.CooMatrix.from_general(n, fi[order], fj[order], fv[order], x)
    mo = orc.read_unsym(n, fi[order], fj[order], fv[order], x)
    xa, ad = m.build_graph(); xo, ao = orc.graph(mo)
    assert np.array_equal(xa, xo) and np.array_equal(ad, ao)
    P, W = 3, 512
    part = rng.integers(0, P, n).astype(np.uint32)
    m.set_plan(P, W, 1)
    m.reorder_with_partition(part)
    ro = orc.reorder(mo, P, W, part)
    a = m.arrays()
    for k in util.COO_KEYS:
        assert np.array_equal(a[k], ro[k]), k
    lay = api.Layout(m)
    st = lay.stats()
    assert st["nnzEll"] + st["nnzRemInSlice"] + st["nnzOverflow"] == st["nnz"]
    e2 = lay.to_reference()
    eo = orc.convert(ro)
    for k in util.EHYB_KEYS:
        assert np.array_equal(e2[k], eo[k]), k
    lay.free(); m.free()


def test_mtmetis_helper_and_matrix_reorder_entry_point(orc):
    """matrixReorder end to end (graph -> pinned mt-metis binary -> permutation) equals the oracle
    fed with the same partition; the partition respects the reference's balance (ubvec 1.001)."""
    if not O.MTMETIS_BIN.exists():
        pytest.skip("bin/ehyb_mtmetis not built")
    kind, dims, P, W = "st27", (24, 24, 24), 6, 2560
    n, li, lj, lv = util.lower_entries(kind, dims)
    m = api.CooMatrix.from_lower(n, li, lj, lv)
    m.set_plan(P, W, 1)
    m.reorder()
    a = m.arrays()
    sizes = np.diff(a["partBoundary"])
    assert sizes.sum() == n and sizes.max() <= 1.05 * n / P
    part = np.empty(n, np.uint32)
    part[:] = np.searchsorted(a["partBoundary"], a["reorderList"], side="right") - 1
    ro = orc.reorder(orc.read_sym(n, li, lj, lv), P, W, part)
    for k in util.COO_KEYS:
        assert np.array_equal(a[k], ro[k]), k
    assert np.array_equal(part, util.metis_partition(kind, dims, P))  # 1 thread: deterministic
    m.free()


def test_plan_auto_picks_by_size_and_density():
    """ehyb_plan_auto (plan.c): one partition per SM for the staged kernel when matrix + vectors fit 3/4 of
    L2 (BASELINE config 1), the persistent kernel's plan up to ~40 entries per row (config 2), the staged
    plan for denser rows (config 3)."""
    dev = api.device_info_b200()
    p1, k1 = api.plan_auto(1048576, 5238784, dev)               # config 1: 5-point 1024^2, 61 MB
    assert k1 == api.KERNEL_STAGED and p1.nParts == dev.sm_count and p1.ctasPerPart == 1
    assert p1.W >= -(-1048576 // dev.sm_count) and p1.W % 64 == 0 and p1.W <= 65472
    p2, k2 = api.plan_auto(2097152, 55742968, dev)              # config 2: the plan bench.py times
    assert k2 == api.KERNEL_PERSISTENT and (p2.nParts, p2.W) == (444, 4864)
    pr = api.plan(2097152, dev, kernel=api.KERNEL_PERSISTENT)
    assert (p2.nParts, p2.W, p2.ctasPerPart) == (pr.nParts, pr.W, pr.ctasPerPart)
    p3, k3 = api.plan_auto(3000000, 238172328, dev)             # config 3: ~79 entries per row
    ps = api.plan(3000000, dev)
    assert k3 == api.KERNEL_STAGED and (p3.nParts, p3.W) == (ps.nParts, ps.W)
    p4, k4 = api.plan_auto(4096, 20000, dev)                    # too small for a partition per SM: the generic small-matrix rule
    assert p4.nParts >= 1 and p4.W >= 64
