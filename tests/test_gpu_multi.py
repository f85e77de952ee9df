"""Multi-GPU parity on real GPUs (needs >= 2; run with `gpurun --gpus 2 -- python -m pytest tests -m gpu`).

One process per GPU runs tests/mg_gpu_worker.py: distributed products through the C ABI with
both exchanges (peer-memory push inside the main kernel, NCCL send/recv) against the oracle."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


CASES = [
    # exchange, level-2 partition, per-rank grid
    ("p2p", "blocks", "10x9x5"),     # tiny: a few partitions, several CTAs each, one pushing CTA
    ("p2p", "metis", "24x20x9"),
    ("p2p", "metis", "64x64x24"),    # ~100 k rows per rank: many partitions, several pushing CTAs
    ("p2p", "oneway", "-"),          # rank r needs x from rank r-1 only: one-way halo dependencies
    ("nccl", "oneway", "-"),
    ("nccl", "metis", "24x20x9"),
    ("nccl", "blocks", "64x64x24"),
    # BASELINE.json config 5's pipeline (GLOBAL grid, bricks, level-1 owners, streamed build, closed-form check)
    ("p2p", "grid-metis", "48x40x36"),
    ("p2p", "grid-scatter", "40x40x24"),    # every rank neighbours every other
    ("nccl", "grid-metis", "48x40x36"),
    ("p2p", "grid-metis", "160x160x96"),    # 2.4 M rows: hundreds of partitions per GPU, many per CTA
]


@pytest.mark.parametrize("exchange,partition,grid", CASES)
@pytest.mark.parametrize("kernel", ["staged", "persistent"])
@pytest.mark.parametrize("world", [2, 4])
def test_distributed_product_on_gpus(world, exchange, partition, grid, kernel):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    if "metis" in partition and not (ROOT / "bin" / "ehyb_mtmetis").exists():
        pytest.skip("bin/ehyb_mtmetis not built")
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="4", EHYB_P2P_TIMEOUT_MS="20000",
                   EHYB_MG_KERNEL={"staged": "2", "persistent": "3"}[kernel])
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "mg_gpu_worker.py"), exchange, partition, grid],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("rank %d ok" % r) in out, "rank %d:\n%s" % (r, out[-3000:])
