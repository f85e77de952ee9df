"""ehyb_pcg_solve on the GPU against the same algorithm in numpy over the
oracle's CSR product: iteration counts agree, the solution is the known one, the reported
residuals are what a CPU evaluation of b - A x gives."""
import numpy as np
import pytest

from ehyb_spmv_gpu_b200 import api
from tests import util

pytestmark = pytest.mark.gpu


def _pcg_numpy(orc, a, b, dinv, rtol, max_iters):
    x = np.zeros_like(b); r = b.copy(); z = dinv * r; p = z.copy()
    rz = float(r @ z); bb = float(b @ b)
    for k in range(max_iters):
        q = orc.csr_spmv(a["rowIdx"], a["J"], a["V"], p)
        alpha = rz / float(p @ q)
        x += alpha * p
        r -= alpha * q
        z = dinv * r
        rz_new = float(r @ z)
        if np.sqrt(float(r @ r) / bb) <= rtol:
            return x, k + 1
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, max_iters


@pytest.mark.parametrize("kind,dims,P,W", [("st27", (24, 24, 24), 6, 2432), ("lap2d", (96, 96), 5, 1920)])
@pytest.mark.parametrize("jacobi", [True, False])
def test_pcg_matches_numpy_pcg(orc, kind, dims, P, W, jacobi):
    n = util.lower_entries(kind, dims)[0]
    m = util.product_pipeline(kind, dims, P, W, 1)
    a = m.arrays()
    lay = api.Layout(m)
    s = api.Session(lay)
    x_true = util.x_random(n, 5)
    b = orc.csr_spmv(a["rowIdx"], a["J"], a["V"], x_true)            # permuted numbering throughout
    diag = np.zeros(n)
    rows = np.repeat(np.arange(n), np.diff(a["rowIdx"]))
    on = a["J"] == rows
    diag[rows[on]] = a["V"][on]
    assert np.all(diag > 0)
    rtol = 1e-10
    x, info = s.pcg_solve(b, diag if jacobi else None, max_iters=2000, rtol=rtol, check_every=1)
    x_ref, it_ref = _pcg_numpy(orc, a, b, 1.0 / diag if jacobi else np.ones(n), rtol, 2000)
    assert info["converged"] and abs(info["iters"] - it_ref) <= 2, (info, it_ref)
    assert np.linalg.norm(x - x_true) <= 1e-7 * np.linalg.norm(x_true)
    assert np.linalg.norm(x - x_ref) <= 1e-7 * np.linalg.norm(x_true)
    true_res = np.linalg.norm(b - orc.csr_spmv(a["rowIdx"], a["J"], a["V"], x)) / np.linalg.norm(b)
    assert true_res <= 20 * rtol and abs(info["true_rel_residual"] - true_res) <= 1e-9 + 0.05 * true_res
    assert info["rel_residual"] <= rtol and info["ms"] > 0
    # the default checking interval only rounds the iteration count up
    x8, info8 = s.pcg_solve(b, diag if jacobi else None, max_iters=2000, rtol=rtol, check_every=8)
    assert info8["converged"] and info["iters"] <= info8["iters"] < info["iters"] + 8 and info8["iters"] % 8 == 0
    assert np.linalg.norm(x8 - x_true) <= 1e-7 * np.linalg.norm(x_true)
    # iteration limit and the trivial right-hand side
    _, lim = s.pcg_solve(b, None, max_iters=5, rtol=rtol)
    assert not lim["converged"] and lim["iters"] == 5 and lim["true_rel_residual"] > rtol
    x0, zero = s.pcg_solve(np.zeros(n), None)
    assert zero["converged"] and zero["iters"] == 0 and not x0.any()
    s.free(); lay.free(); m.free()


@pytest.mark.parametrize("kernel", ["persistent", "staged"])
def test_product_with_fused_dot(orc, kernel):
    """ehyb_spmv_dot: y bit-identical to the plain product, x . y equal to the host's dot product of the same
    vectors (the p.Ap of a CG iteration is taken from the x window while y is stored)."""
    kind, dims = "st27", (40, 40, 40)
    n = util.lower_entries(kind, dims)[0]
    P, W = (450, 192) if kernel == "persistent" else (20, 3264)
    m = util.product_pipeline(kind, dims, P, W, 1)
    lay = api.Layout(m)
    s = api.Session(lay, kernel=api.KERNEL_PERSISTENT if kernel == "persistent" else api.KERNEL_STAGED)
    assert s.kernel_name() == ("ehyb_persistent_kernel" if kernel == "persistent" else "ehyb_staged_kernel")
    assert s.spmv_dot_supported()
    for seed in (1, 2, 3):
        x = util.x_random(n, seed)
        y_plain = s.spmv_host(x)
        y, d = s.spmv_dot_host(x)
        assert np.array_equal(y, y_plain), "the fused-dot build of the kernel computes another y"
        ref = float(np.dot(x, y_plain))
        assert abs(d - ref) <= 1e-12 * float(np.dot(np.abs(x), np.abs(y_plain))), (d, ref)
    s.free(); lay.free(); m.free()


def test_pcg_on_another_device_than_the_current_one(orc):
    """the solver allocates and launches on the SESSION's device whatever the caller's
    current device is, and restores it."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    kind, dims = "lap2d", (64, 64)
    n = util.lower_entries(kind, dims)[0]
    m = util.product_pipeline(kind, dims, 4, 1088, 1)
    a = m.arrays()
    lay = api.Layout(m)
    s = api.Session(lay, device=1)
    torch.cuda.set_device(0)
    x_true = util.x_random(n, 5)
    b = orc.csr_spmv(a["rowIdx"], a["J"], a["V"], x_true)
    x, info = s.pcg_solve(b, None, max_iters=2000, rtol=1e-10)
    assert info["converged"] and np.linalg.norm(x - x_true) <= 1e-7 * np.linalg.norm(x_true)
    assert torch.cuda.current_device() == 0
    s.free(); lay.free(); m.free()
