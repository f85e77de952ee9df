import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("EHYB_MTMETIS_BIN", str(ROOT / "bin" / "ehyb_mtmetis"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O
    return O.Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle as O
    if not O.Reference.available():
        pytest.skip("oracle/_ref/libehyb_ref.so not built (needs /root/reference at build time)")
    return O.Reference()


@pytest.fixture(scope="session")
def lib():
    from ehyb_spmv_gpu_b200 import _lib
    return _lib.load()
