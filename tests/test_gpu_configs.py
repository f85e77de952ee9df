"""BASELINE.json configs 2 (at the plan bench.py times), 3 and 4 at their full sizes on the GPU, and the
reference's own session entry points called the way its driver calls them.

  * config 2 at the PERSISTENT plan (P = 444, W = 4 864: the headline kernel at the headline plan): A*1
    equals the closed-form integer row sums bit for bit, the driver's x inside the gate;
  * config 3, elasticity 100^3 x 3 dof (n 3 000 000, nnz 238 172 328): the oracle's CSR gate, A*1 exact
    (every entry and every partial sum is a multiple of 0.25 below 2^53: fp64 is exact in any order),
    linearity;
  * config 4, R-MAT scale 24 (n 16 777 216, 2^28 generated edges, duplicates summed) through the
    general (unsymmetric) path: the oracle's CSR gate.  The reference itself cannot run this input at
    any scale, so the CPU CSR product is the only oracle;
  * spmvGPuEHYB(matrixCOO*, x, y, MAXIter, &realIter) (reference spmv.h:75-78) and ehyb_upload +
    ehyb_describe + matrixVectorEHYB / matrixVectorEHYB_small (reference kernel.h:52-60) with plain
    device pointers.
"""
import ctypes as C

import numpy as np
import pytest

from ehyb_spmv_gpu_b200 import _lib as L
from ehyb_spmv_gpu_b200 import api
from tests import util
from tests.test_gpu_fullsize import _row_sums

pytestmark = pytest.mark.gpu


def test_config2_persistent_plan(orc):
    dims = (128, 128, 128)
    n, li, lj, lv = api.gen_lower(api.GEN_STENCIL27, *dims)
    x = api.x_reference(n)
    m = api.CooMatrix.from_lower(n, li, lj, lv, x)
    del li, lj, lv
    pl = api.plan(n, api.device_query(0), kernel=api.KERNEL_PERSISTENT)
    assert (pl.nParts, pl.W) == (444, 4864)          # what bench.py times
    m.set_plan(pl.nParts, pl.W, pl.ctasPerPart)
    m.reorder()
    lay = api.Layout(m)
    st = lay.stats()
    s = api.Session(lay)
    assert s.kernel_name() == "ehyb_persistent_kernel"
    a = m.arrays()
    y1 = m.vector_recover(s.spmv_host(m.vector_reorder(np.ones(n))))
    assert np.array_equal(y1, _row_sums(api.GEN_STENCIL27, dims)), "A*1 differs from the closed-form row sums"
    xr = m.vector_reorder(x)
    y = s.spmv_host(xr)
    y_ref = orc.csr_spmv(a["rowIdx"], a["J"], a["V"], xr)
    absAx = orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], xr)
    util.assert_within_gate(y, y_ref, absAx)
    util.assert_within_gate(m.vector_recover(y), m.y_golden, m.vector_recover(absAx))
    # back to back (programmatic dependent launch) the same bits come out, and the staged kernel agrees
    s.set_x(xr)
    s.time_spmv(2, 20)
    assert st["nOverflow"] == 0 and np.array_equal(s.get_y(), y)
    s2 = api.Session(lay, kernel=api.KERNEL_STAGED)
    assert np.array_equal(s2.spmv_host(xr), y), "persistent and staged kernels differ at config 2"
    s.free(); s2.free(); lay.free(); m.free()


def test_config3_elasticity_full_size(orc):
    dims = (100, 100, 100)
    n, li, lj, lv = api.gen_lower(api.GEN_ELASTICITY, *dims)
    assert n == 3_000_000
    x = api.x_reference(n)
    m = api.CooMatrix.from_lower(n, li, lj, lv, x)
    del li, lj, lv
    assert m.nnz == 238_172_328                     # 298^3 * 9, SURVEY.md Appendix D
    pl = api.plan(n, api.device_query(0))
    m.set_plan(pl.nParts, pl.W, pl.ctasPerPart)
    m.reorder()                                     # pinned mt-metis, one thread: ~30 s
    lay = api.Layout(m)
    st = lay.stats()
    assert st["nnzEll"] + st["nnzRemInSlice"] + st["nnzOverflow"] == st["nnz"] == m.nnz
    s = api.Session(lay)
    a = m.arrays()
    xr = m.vector_reorder(x)
    y = s.spmv_host(xr)
    y_ref = orc.csr_spmv(a["rowIdx"], a["J"], a["V"], xr)
    absAx = orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], xr)
    util.assert_within_gate(y, y_ref, absAx)
    util.assert_within_gate(m.vector_recover(y), m.y_golden, m.vector_recover(absAx))
    # A*1: values are 100 and -(1 + 0.25 k): every partial sum is a multiple of 0.25, exact in any order
    ones = np.ones(n)
    y1 = s.spmv_host(ones)
    assert np.array_equal(y1, orc.csr_spmv(a["rowIdx"], a["J"], a["V"], ones)), "A*1 is not exact"
    # linearity
    z = m.vector_reorder(util.x_random(n, 3))
    yz = s.spmv_host(z)
    bound = 2.0 * absAx + orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], z)
    util.assert_within_gate(s.spmv_host(2.0 * xr + z), 2.0 * y + yz, 2.0 * bound)
    s.free(); lay.free(); m.free()


@pytest.mark.parametrize("scale", [24])
def test_config4_rmat_full_size(orc, scale):
    n, fi, fj, fv = api.gen_rmat(scale, 16, seed=1, add_diagonal=False)
    assert n == 1 << scale
    x = np.random.default_rng(0).uniform(-0.1, 0.1, n)
    m = api.CooMatrix.from_general(n, fi, fj, fv, x)
    del fi, fj, fv
    pl = api.plan(n, api.device_query(0))
    m.set_plan(pl.nParts, pl.W, pl.ctasPerPart)
    # contiguous row blocks as the partition: mt-metis needs many minutes on a power-law graph of this size,
    # and the format decision (layout.c: coverage below 20 % -> everything in the row-sorted list) makes the
    # partition irrelevant for the product; rows are still sorted by in-partition count inside the blocks
    m.reorder_with_partition((np.arange(n, dtype=np.int64) * pl.nParts // n).astype(np.uint32))
    lay = api.Layout(m)
    st = lay.stats()
    assert st["nnzEll"] + st["nnzRemInSlice"] + st["nnzOverflow"] == st["nnz"] == m.nnz
    s = api.Session(lay)
    assert s.kernel_name() == "ehyb_ovfstream_kernel"   # the CSR-like stream, not the COO list (no atomics)
    a = m.arrays()
    xr = m.vector_reorder(x)
    y = s.spmv_host(xr)
    y_ref = orc.csr_spmv(a["rowIdx"], a["J"], a["V"], xr)
    absAx = orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], xr)
    util.assert_within_gate(y, y_ref, absAx)
    util.assert_within_gate(m.vector_recover(y), m.y_golden, m.vector_recover(absAx))
    assert np.array_equal(s.spmv_host(xr), y), "the stream kernel is deterministic: a second product must give the same bits"
    # a second product of another x: nothing stale from the first, still inside the gate
    z = m.vector_reorder(util.x_random(n, 5))
    yz = s.spmv_host(z)
    util.assert_within_gate(yz, orc.csr_spmv(a["rowIdx"], a["J"], a["V"], z), orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], z))
    s.free(); lay.free(); m.free()


def _small_matrix(orc):
    kind, dims, P, W = "st27", (32, 32, 32), 8, 4224
    n = util.lower_entries(kind, dims)[0]
    x = orc.x_reference(n)
    m = util.product_pipeline(kind, dims, P, W, 1, x=x)
    return m, x


def test_spmvGPuEHYB_reference_signature(orc, lib, capfd):
    """void spmvGPuEHYB(matrixCOO*, const double*, double*, const int MAXIter, int* realIter), called like
    reference solver_test.c:382 on a reordered matrixCOO: y comes back in the permuted numbering, the log
    lines of the reference are printed, realIter (never written by the reference) is the iteration count."""
    m, x = _small_matrix(orc)
    xr = m.vector_reorder(x)
    y = np.full(m.n, np.nan)
    real = C.c_int(-1)
    lib.spmvGPuEHYB.restype = None
    lib.spmvGPuEHYB(C.byref(m.c), xr.ctypes.data_as(L.c_dbl_p), y.ctypes.data_as(L.c_dbl_p), 5, C.byref(real))
    out = capfd.readouterr().out
    assert real.value == 5
    for line in ("toER is", "wasteElement is", "sizeER is", "iter is 5, time is", "GPU Gflops is"):
        assert line in out, "log line '%s' missing:\n%s" % (line, out)
    a = m.arrays()
    util.assert_within_gate(y, orc.csr_spmv(a["rowIdx"], a["J"], a["V"], xr), orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], xr))
    util.assert_within_gate(m.vector_recover(y), m.y_golden, m.vector_recover(orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], xr)))
    m.free()


@pytest.mark.parametrize("entry", ["matrixVectorEHYB", "matrixVectorEHYB_small"])
def test_matrixVectorEHYB_device_pointers(orc, lib, entry):
    """The per-product launchers of reference kernel.h:52-60 on a matrixEHYB that describes an uploaded
    session (ehyb_upload + ehyb_describe), with device vectors allocated by somebody else (cudaMalloc
    through torch): x changes between two products, both are checked."""
    import torch
    m, x = _small_matrix(orc)
    lay = api.Layout(m)
    h = C.c_void_p()
    L.check(lib, lib.ehyb_upload(lay.h, None, C.byref(h)), "ehyb_upload")
    d = L.MatrixEHYB()
    L.check(lib, lib.ehyb_describe(h, C.byref(d)), "ehyb_describe")
    assert d.dimension == m.n and d.nParts == 8 and d.b200 == h.value
    a = m.arrays()
    getattr(lib, entry).restype = None
    for seed in (0, 1):
        xr = m.vector_reorder(x if seed == 0 else util.x_random(m.n, seed))
        xd = torch.from_numpy(xr).cuda()
        yd = torch.full((m.n,), float("nan"), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        if entry == "matrixVectorEHYB":
            lib.matrixVectorEHYB(C.byref(d), C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr()))
        else:
            bias = torch.zeros(64, dtype=torch.int32, device="cuda")   # the reference's work counters: accepted, unused
            lib.matrixVectorEHYB_small(C.byref(d), C.c_void_p(bias.data_ptr()), C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr()))
        L.check(lib, lib.ehyb_sync(h), "ehyb_sync")
        y = yd.cpu().numpy()
        util.assert_within_gate(y, orc.csr_spmv(a["rowIdx"], a["J"], a["V"], xr), orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], xr))
    lib.ehyb_free(h)
    lay.free(); m.free()
