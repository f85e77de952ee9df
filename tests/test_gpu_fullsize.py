"""Parity at BASELINE.json's full sizes (configs 1 and 2) on the GPU, through properties that do
not depend on the size: A*1 is exact in fp64 for these stencils (integer row sums known in closed
form) - a bit-exact check of every row without the oracle -, linearity, and the oracle's CSR
product inside the accuracy gate.  The whole product pipeline runs: generator -> reader
expansion -> plan from the device -> pinned mt-metis -> reorder -> tuned layout -> session."""
import numpy as np
import pytest

from ehyb_spmv_gpu_b200 import api
from oracle import oracle as O
from tests import util

pytestmark = pytest.mark.gpu


def _row_sums(kind, dims):
    """sum_j A[i, j] of the generated stencils: diagonal minus the number of neighbours."""
    if kind == api.GEN_LAPLACE2D:
        nx, ny = dims
        cx = 1 + (np.arange(nx) > 0) + (np.arange(nx) < nx - 1)
        cy = 1 + (np.arange(ny) > 0) + (np.arange(ny) < ny - 1)
        # 5-point: neighbours = (cx - 1) + (cy - 1)
        return (4.0 - ((cx - 1)[None, :] + (cy - 1)[:, None])).reshape(-1)
    nx, ny, nz = dims
    cx = 1 + (np.arange(nx) > 0) + (np.arange(nx) < nx - 1)
    cy = 1 + (np.arange(ny) > 0) + (np.arange(ny) < ny - 1)
    cz = 1 + (np.arange(nz) > 0) + (np.arange(nz) < nz - 1)
    cnt = cz[:, None, None] * cy[None, :, None] * cx[None, None, :]
    return (26.0 - (cnt - 1)).reshape(-1).astype(np.float64)


@pytest.mark.parametrize("name,kind,dims", [("config 1: 5-point 1024^2", api.GEN_LAPLACE2D, (1024, 1024)),
                                            ("config 2: 27-point 128^3", api.GEN_STENCIL27, (128, 128, 128))])
def test_full_size_properties(orc, name, kind, dims):
    n, li, lj, lv = api.gen_lower(kind, *dims)
    x = api.x_reference(n)
    m = api.CooMatrix.from_lower(n, li, lj, lv, x)
    del li, lj, lv
    pl = api.plan(n, api.device_query(0))
    m.set_plan(pl.nParts, pl.W, pl.ctasPerPart)
    m.reorder()                      # pinned mt-metis through bin/ehyb_mtmetis
    lay = api.Layout(m)
    st = lay.stats()
    assert st["nnzEll"] + st["nnzRemInSlice"] + st["nnzOverflow"] == st["nnz"] == m.nnz
    s = api.Session(lay)
    a = m.arrays()

    # 1. A * 1: every partial sum is a small integer, so any summation order is exact
    ones = np.ones(n)
    y1 = m.vector_recover(s.spmv_host(m.vector_reorder(ones)))
    assert np.array_equal(y1, _row_sums(kind, dims)), "A*1 differs from the closed-form row sums"

    # 2. the driver's x: inside the gate of the oracle's CSR product and of the golden y
    xr = m.vector_reorder(x)
    y = s.spmv_host(xr)
    y_ref = orc.csr_spmv(a["rowIdx"], a["J"], a["V"], xr)
    absAx = orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], xr)
    util.assert_within_gate(y, y_ref, absAx)
    util.assert_within_gate(m.vector_recover(y), m.y_golden, m.vector_recover(absAx))

    # 3. linearity: A(2x + z) = 2 A x + A z  (scaling by 2 is exact; the rest inside the gate)
    z = util.x_random(n, 11)
    zr = m.vector_reorder(z)
    yz = s.spmv_host(zr)
    ycomb = s.spmv_host(2.0 * xr + zr)
    bound = 2.0 * absAx + orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], zr)
    util.assert_within_gate(ycomb, 2.0 * y + yz, 2.0 * bound)

    # 4. a second product of the same x is bit-identical when nothing goes through the overflow list
    if st["nOverflow"] == 0:
        assert np.array_equal(s.spmv_host(xr), y)
    s.free(); lay.free(); m.free()
