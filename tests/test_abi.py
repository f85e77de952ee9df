"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

from ehyb_spmv_gpu_b200 import _lib

INC = Path(__file__).resolve().parent.parent / "include"


CUSPARSE_HEADER = "ehyb_cusparse.h"  # its symbols live in libehyb_cusparse.so


def declared_functions(only=None):
    names = set()
    for h in INC.glob("*.h"):
        if (h.name == CUSPARSE_HEADER) != (only == CUSPARSE_HEADER):
            continue
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
        text = re.sub(r"static inline[^{]*\{.*?\n\}", "", text, flags=re.S)
        text = re.sub(r"typedef\s+\w+\s*\(\*\w+\)\s*\([^;]*\);", "", text, flags=re.S)
        for mt in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text):
            name = mt.group(1)
            if name not in ("defined", "sizeof", "void"):
                names.add(name)
    return names


def test_every_declared_symbol_is_exported(lib):
    decl = declared_functions()
    assert {"spmvGPuEHYB", "matrixVectorEHYB", "matrixVectorEHYB_small", "COO2EHYB", "matrixReorder",
            "matrixReorder_unsym", "vectorReorder", "vectorRecover", "mm_read_banner", "ehyb_upload",
            "ehyb_spmv", "ehyb_layout_build"} <= decl
    missing = sorted(s for s in decl if not hasattr(lib, s))
    assert not missing, missing
    assert set(_lib.EXPORTS) <= decl


def test_comparison_library_exports_its_header():
    """libehyb_cusparse.so (spmvGeneric, the reference's cuSPARSE comparison) loads on a box
    without a GPU and exports what include/ehyb_cusparse.h declares."""
    path = _lib.PKG / "lib" / "libehyb_cusparse.so"
    assert path.exists(), "build it with __graft_entry__.build()"
    lib = C.CDLL(str(path))
    decl = declared_functions(only=CUSPARSE_HEADER)
    assert {"spmvGeneric", "ehyb_cusparse_spmv"} <= decl
    assert not [s for s in decl if not hasattr(lib, s)]


def test_struct_layouts_match_the_reference(ref):
    """matrixCOO is the reference's struct; matrixEHYB is the reference's plus one appended pointer."""
    assert C.sizeof(_lib.MatrixCOO) == ref.L.ref_sizeof_matrixCOO()
    assert C.sizeof(_lib.MatrixEHYB) == ref.L.ref_sizeof_matrixEHYB() + C.sizeof(C.c_void_p)
    from oracle.oracle import MatrixEHYB as RefEHYB
    for (name, _), (rname, _) in zip(_lib.MatrixEHYB._fields_, RefEHYB._fields_):
        assert name == rname
        assert getattr(_lib.MatrixEHYB, name).offset == getattr(RefEHYB, rname).offset


def test_version_plan_and_errors(lib):
    from ehyb_spmv_gpu_b200 import api
    assert b"sm_100a" in lib.ehyb_version()
    d = api.device_info_b200()
    assert (d.sm_count, d.smem_optin_bytes) == (148, 232448)
    p = api.plan(2097152, d)
    assert p.nParts % 148 == 0 and p.W % 64 == 0 and p.W * 8 <= 232448 and p.W >= 2097152 / p.nParts
    small = api.plan(20000, d)
    assert small.nParts * small.ctasPerPart >= 148
    rc = lib.ehyb_plan(0, C.byref(d), C.byref(p))
    assert rc == -1 and b"ehyb_plan" in lib.ehyb_last_error()
    # the plan for the persistent kernel: three or more partitions per SM, windows small enough that
    # two {window, 24 KB cache} buffers leave 16 warps of staging; small matrices keep the staged plan
    for n in (1048576, 2097152, 16777216):
        st, pe = api.plan(n, d), api.plan(n, d, kernel=api.KERNEL_PERSISTENT)
        assert pe.nParts % 148 == 0 and pe.nParts >= 3 * 148 and pe.nParts > st.nParts and pe.ctasPerPart == 1
        assert pe.W % 64 == 0 and pe.W >= n / pe.nParts and 2 * (pe.W * 8 + 24 * 1024) + 16 * 2 * 2560 + 1664 <= d.smem_optin_bytes
        assert (pe.nParts + 147) // 148 <= 32
    sp = api.plan(20000, d, kernel=api.KERNEL_PERSISTENT)
    assert (sp.nParts, sp.W, sp.ctasPerPart) == (small.nParts, small.W, small.ctasPerPart)


def test_no_gpu_means_loud_failure(lib):
    """Without a CUDA device the session entry point must fail, not fall back."""
    cnt = C.c_int(0)
    rc = lib.ehyb_device_count(C.byref(cnt))
    if rc == 0 and cnt.value > 0:
        return  # on the GPU box this is covered by the gpu tests
    from ehyb_spmv_gpu_b200 import api
    import pytest
    n, li, lj, lv = api.gen_lower(api.GEN_LAPLACE2D, 16, 16)
    m = api.CooMatrix.from_lower(n, li, lj, lv)
    m.set_plan(2, 192, 1)
    m.reorder_with_partition([0] * 128 + [1] * 128)
    lay = api.Layout(m)
    with pytest.raises(_lib.EhybError):
        api.Session(lay)
