"""Binary cache of the finished pipeline (csrc/host/cache.c): a cached layout
is bit-identical to the built one, maps back onto the reference arrays the same way, and a cache
that does not belong to the source file / parameters, or is damaged, is rejected."""
import os

import ctypes as C

import numpy as np
import pytest

from ehyb_spmv_gpu_b200 import _lib, api
from tests import util


def _pipeline(x=None):
    kind, dims, P, W = "st27", (14, 13, 12), 5, 576
    m = util.product_pipeline(kind, dims, P, W, 1, x=x)
    return m, api.Layout(m, er_fill=0.5, cache_cap=48)   # small cache: overflow entries exist too


def _same(a, b):
    ra, rb = a.raw(), b.raw()
    assert a.stats() == b.stats()
    for k in ra:
        assert np.array_equal(ra[k], rb[k]), k


def test_layout_roundtrip_is_bit_identical(tmp_path):
    m, lay = _pipeline()
    assert lay.stats()["nOverflow"] > 0
    path = tmp_path / "m.ehyb"
    lay.save(path)
    back = api.Layout.load(path)
    _same(lay, back)
    # the de-interleave to the reference layout (host bookkeeping arrays travel too)
    ea, eb = lay.to_reference(), back.to_reference()
    for k in util.EHYB_KEYS:
        assert np.array_equal(ea[k], eb[k]), k
    back.free(); lay.free(); m.free()


def test_pipeline_cache_and_rejections(tmp_path):
    n = util.lower_entries("st27", (14, 13, 12))[0]
    x = util.x_random(n, 3)
    m, lay = _pipeline(x)
    a = m.arrays()
    src = tmp_path / "m.mtx"
    src.write_text("%%MatrixMarket matrix coordinate real symmetric\n1 1 1\n1 1 1.0\n")
    absAx = np.abs(m.y_golden) + 1.0
    path = tmp_path / "m.mtx.ehyb"
    lay.cache_save(path, src, True, a["reorderList"], x, m.y_golden, absAx)
    back, vec = api.Layout.cache_load(path, src)
    _same(lay, back)
    assert vec["n"] == n and vec["symmetric"]
    assert np.array_equal(vec["reorderList"], a["reorderList"]) and np.array_equal(vec["x"], x)
    assert np.array_equal(vec["y_golden"], m.y_golden) and np.array_equal(vec["absAx"], absAx)
    # the product evaluated from the cached arrays
    assert np.allclose(util.layout_spmv(back.raw(), m.vector_reorder(x)), m.vector_reorder(m.y_golden), rtol=0, atol=1e-12)
    back.free()
    # a cache written under other layout / partitioner options (the caller's tag) is rejected, not reused
    lib = _lib.load()
    lib.ehyb_cache_set_options_tag(C.c_uint64(0x1234))
    with pytest.raises(_lib.EhybError, match="other layout / partitioner options"):
        api.Layout.cache_load(path, src)
    tagged = tmp_path / "tagged.ehyb"
    lay.cache_save(tagged, src, True, a["reorderList"], x, m.y_golden, absAx)
    again, _ = api.Layout.cache_load(tagged, src)
    again.free()
    lib.ehyb_cache_set_options_tag(C.c_uint64(0))
    with pytest.raises(_lib.EhybError, match="other layout / partitioner options"):
        api.Layout.cache_load(tagged, src)
    # same parameters accepted, other parameters rejected
    st = lay.stats()
    ok = _lib.Plan(st["nParts"], st["W"], st["ctasPerPart"], 0, 1)
    api.Layout.cache_load(path, src, ok)[0].free()
    with pytest.raises(_lib.EhybError, match="holds P="):
        api.Layout.cache_load(path, src, _lib.Plan(st["nParts"] + 1, st["W"], st["ctasPerPart"], 0, 1))
    # the source file changed after the cache was written
    os.utime(src, ns=(1, 1))
    with pytest.raises(_lib.EhybError, match="another version"):
        api.Layout.cache_load(path, src)
    os.utime(src)
    # damage: one flipped byte in the payload, truncation, wrong magic
    raw = bytearray(path.read_bytes())
    flipped = tmp_path / "flipped.ehyb"
    raw2 = bytearray(raw); raw2[len(raw2) // 2] ^= 0x40
    flipped.write_bytes(raw2)
    with pytest.raises(_lib.EhybError, match="checksum|inconsistent"):
        api.Layout.load(flipped)
    (tmp_path / "short.ehyb").write_bytes(raw[: len(raw) - 4096])
    with pytest.raises(_lib.EhybError, match="truncated"):
        api.Layout.load(tmp_path / "short.ehyb")
    raw3 = bytearray(raw); raw3[0] ^= 0xFF
    (tmp_path / "magic.ehyb").write_bytes(raw3)
    with pytest.raises(_lib.EhybError, match="not an EHYB cache"):
        api.Layout.load(tmp_path / "magic.ehyb")
    with pytest.raises(_lib.EhybError, match="not found"):
        api.Layout.load(tmp_path / "absent.ehyb")
    lay.free(); m.free()
