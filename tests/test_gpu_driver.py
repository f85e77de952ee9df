"""bin/spmv.out (the reference's driver workflow, solver_test.c) on the GPU: -i/-m style run
from a .mtx file, exit code from the accuracy gate, and the binary cache - the second run of the
same file must load <file>.ehyb, skip reader / mt-metis / reorder / format build and produce the
same verdict."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest

from ehyb_spmv_gpu_b200 import _lib as L
from ehyb_spmv_gpu_b200 import api

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _run(*args):
    p = subprocess.run([str(ROOT / "bin" / "spmv.out"), *args], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600)
    return p.returncode, p.stdout


def test_driver_from_mtx_and_cache(tmp_path):
    if not (ROOT / "bin" / "spmv.out").exists():
        pytest.skip("bin/spmv.out not built")
    n, li, lj, lv = api.gen_lower(api.GEN_STENCIL27, 40, 40, 40)
    lib = L.load()
    mtx = tmp_path / "st27_40.mtx"
    L.check(lib, lib.ehyb_write_mtx(str(mtx).encode(), n, C.c_int64(len(li)), li.ctypes.data_as(L.c_int_p), lj.ctypes.data_as(L.c_int_p),
                                    lv.ctypes.data_as(L.c_dbl_p), 1), "ehyb_write_mtx")
    cache = Path(str(mtx) + ".ehyb")
    rc, out = _run("-i", "20", "-M", str(mtx))
    assert rc == 0, out[-2000:]
    assert "read symmetric matrix" in out and "cache: wrote" in out and "host stages:" in out and cache.exists()
    assert "0 of %d rows fail" % n in out
    first = [l for l in out.splitlines() if l.startswith("EHYB-B200 bytes")]
    rc, out2 = _run("-i", "20", "-M", str(mtx))
    assert rc == 0, out2[-2000:]
    assert "cache: loaded" in out2 and "start k-way partition" not in out2 and "read symmetric matrix" not in out2
    assert "0 of %d rows fail" % n in out2
    assert [l for l in out2.splitlines() if l.startswith("EHYB-B200 bytes")] == first  # the same layout
    # other parameters: the cache is not used (and rewritten for them); -C: neither read nor written
    rc, out3 = _run("-i", "5", "-M", str(mtx), "-P", "40", "-W", "2048", "-K", "4")
    assert rc == 0 and "rebuilding" in out3 and "cache: wrote" in out3, out3[-2000:]
    stamp = cache.stat().st_mtime_ns
    rc, out4 = _run("-i", "5", "-M", str(mtx), "-C")
    assert rc == 0 and "cache:" not in out4 and cache.stat().st_mtime_ns == stamp, out4[-2000:]
    # a damaged cache is reported and ignored, not trusted
    raw = bytearray(cache.read_bytes()); raw[len(raw) // 2] ^= 1
    cache.write_bytes(raw)
    rc, out5 = _run("-i", "5", "-M", str(mtx), "-P", "40", "-W", "2048", "-K", "4")
    assert rc == 0 and "cache not used" in out5 and "start k-way partition" in out5, out5[-2000:]


@pytest.mark.parametrize("gpus", [1, 2, 4])
def test_driver_multi_gpu_mode(gpus):
    """./spmv.out -G <gpus> -g st27:NX:NY:NZ: BASELINE.json config 5's pipeline from the C driver alone -
    one process, one host thread per GPU, level 1 by mt-metis on the brick graph, streamed format build,
    halo exchange inside the persistent kernel over plain peer access; per-GPU lines, the reference's
    `iter is ...` line for the whole job, exit code from the accuracy gate over every row."""
    import torch
    if not (ROOT / "bin" / "spmv.out").exists():
        pytest.skip("bin/spmv.out not built")
    if torch.cuda.device_count() < gpus:
        pytest.skip("needs %d GPUs" % gpus)
    dims = (96, 80, 64)
    rc, out = _run("-G", str(gpus), "-g", "st27:%d:%d:%d" % dims, "-B", "16x16x8", "-i", "20")
    assert rc == 0, out[-3000:]
    n = dims[0] * dims[1] * dims[2]
    assert "0 of %d rows fail" % n in out, out[-3000:]
    for g in range(gpus):
        assert "GPU %d: " % g in out
    assert "iter is 20, time is" in out and "EHYB-B200 multi-GPU: %d GPUs" % gpus in out
    assert out.count("gate 0 rows fail") == gpus
