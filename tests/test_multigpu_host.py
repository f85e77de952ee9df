"""N > 1 path on CPU: world_size-2 and -3 gloo jobs run tests/mg_worker.py (host logic of the
multi-GPU layer + emulated halo exchange, checked against the global product)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,partition,exchange", [(2, "blocks", "nccl"), (3, "blocks", "p2p"), (2, "metis", "p2p"),
                                                      (2, "metis", "nccl"), (3, "oneway", "p2p"), (3, "oneway", "nccl")])
def test_distributed_product_gloo(world, partition, exchange):
    if partition == "metis" and not (ROOT / "bin" / "ehyb_mtmetis").exists():
        pytest.skip("bin/ehyb_mtmetis not built")
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "mg_worker.py"), partition, exchange], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("rank %d ok" % r) in out, out[-3000:]
