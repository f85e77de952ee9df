"""Worker of tests/test_multigpu_host.py: one rank of a world_size-N gloo job on CPU.

Exercises the host side of the multi-GPU layer (ehyb_mg_local_*: halo, send lists, level-2
reorder with halo columns, layout for either exchange) and emulates the per-product exchange
with gloo send/recv - for the peer-memory exchange with the very destination offsets the
device path stores to - checking the distributed product, evaluated from the device-facing
layout arrays, against the global CSR product."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("EHYB_MTMETIS_BIN", str(ROOT / "bin" / "ehyb_mtmetis"))


def main():
    import torch
    import torch.distributed as dist
    from ehyb_spmv_gpu_b200 import multigpu as mg
    from oracle import oracle as O

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    partition = sys.argv[1] if len(sys.argv) > 1 else "blocks"
    exchange = sys.argv[2] if len(sys.argv) > 2 else "nccl"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid = (24, 20, 9) if partition == "metis" else (10, 9, 5)  # per rank
    if partition == "oneway":
        # random block lower-bidiagonal matrix: rank r needs x from rank r-1 only (one-way halo)
        sys.path.insert(0, str(ROOT / "tests"))
        from mg_gpu_worker import setup_oneway
        blk, rowStarts = setup_oneway(rank, world, dist, exchange, n_local=1500)
    else:
        blk, rowStarts = mg.setup_slab(rank, world, grid, dist, partition, exchange)
    orc = O.Oracle()

    # --- structural checks -------------------------------------------------------------
    N = int(rowStarts[-1])
    r0, r1 = int(rowStarts[rank]), int(rowStarts[rank + 1])
    if partition != "oneway":
        plane = grid[0] * grid[1]
        expect_halo = plane * ((rank > 0) + (rank < world - 1))
        assert blk.nHalo == expect_halo, (blk.nHalo, expect_halo)
    assert np.all((blk.haloGlobal < r0) | (blk.haloGlobal >= r1)) and np.all(np.diff(blk.haloGlobal) > 0)
    assert blk.stats["ncols"] == blk.n + blk.nHalo
    J = blk.coo["J"]
    from ehyb_spmv_gpu_b200 import api
    lay = api.Layout.__new__(api.Layout)   # view of the block's layout (owned by the block)
    lay.lib, lay.h, lay.v = blk.lib, None, api.LayoutView()
    api.check(blk.lib, blk.lib.ehyb_layout_get(blk.layout, api.C.byref(lay.v)), "ehyb_layout_get")
    raw = lay.raw()
    if exchange == "nccl":
        # every halo entry sits in the overflow list, none in a remainder cache
        assert blk.stats["haloInOverflow"] == (1 if blk.nHalo else 0)   # (a rank without halo columns has nothing to move)
        assert blk.stats["nOverflow"] >= int(np.count_nonzero(J >= blk.n))
        assert not np.any(raw["cacheCols"] >= blk.n)
    else:
        # halo columns are ordinary remainder columns: they show up in the caches (or overflow)
        assert blk.stats["haloInOverflow"] == 0
        if blk.nHalo:
            assert np.any(raw["cacheCols"] >= blk.n) or np.any(raw["ovfCol"] >= blk.n)
        assert int(raw["cacheCols"].max(initial=0)) < blk.n + blk.nHalo
    assert blk.stats["nnzEll"] + blk.stats["nnzRemInSlice"] + blk.stats["nnzOverflow"] == blk.stats["nnz"]

    # --- one emulated product ---------------------------------------------------------
    x_nat = mg.x_of_global(np.arange(r0, r1))
    x_perm = np.empty(blk.n)
    x_perm[blk.coo["reorderList"]] = x_nat
    packed = x_perm[blk.sendIdx]                        # what the pack kernel / the push gathers
    halo = np.full(blk.nHalo, np.nan)
    reqs, so, ro = [], 0, 0
    send_bufs = []
    if exchange == "p2p":
        # the push stores entry k of the list for peer g at halo_g[recvOffsetOnPeer[g] + k]
        gathered = [None] * world
        dist.all_gather_object(gathered, [int(c) for c in blk.recvCount])
        off = blk.recv_offsets_on_peers(gathered)
        mine = np.concatenate([[0], np.cumsum(blk.recvCount)])
    for g in range(world):
        sc, rc = int(blk.sendCount[g]), int(blk.recvCount[g])
        if sc:
            t = torch.from_numpy(np.ascontiguousarray(packed[so:so + sc]))
            send_bufs.append(t)
            reqs.append(dist.isend(t, g))
            if exchange == "p2p":
                o = torch.tensor([int(off[g])], dtype=torch.int64)
                send_bufs.append(o)
                reqs.append(dist.isend(o, g, tag=1))
        if rc:
            t = torch.from_numpy(halo[ro:ro + rc])
            reqs.append(dist.irecv(t, g))
            if exchange == "p2p":
                o = torch.zeros(1, dtype=torch.int64)
                dist.recv(o, g, tag=1)
                assert int(o.item()) == int(mine[g]) == ro, "peer %d would store at %d, its entries live at %d" % (g, int(o.item()), ro)
        so += sc
        ro += rc
    for q in reqs:
        q.wait()
    assert np.array_equal(halo, mg.x_of_global(blk.haloGlobal)), "halo exchange delivered the wrong entries"
    x_ext = np.concatenate([x_perm, halo])
    y_perm = orc.csr_spmv(blk.coo["rowIdx"], blk.coo["J"], blk.coo["V"], x_ext)
    # the same product from the device-facing layout arrays (what the kernels read)
    from tests import util
    y_lay = util.layout_spmv(raw, x_ext)
    a_perm = orc.csr_abs_spmv(blk.coo["rowIdx"], blk.coo["J"], blk.coo["V"], x_ext)
    assert np.all(np.abs(y_lay - y_perm) <= 1e-12 * a_perm), np.abs(y_lay - y_perm).max()
    y_nat = y_perm[blk.coo["reorderList"]]

    xg = mg.x_of_global(np.arange(N))
    if partition == "oneway":
        # reference: this rank's rows (global columns, natural order) times the global x
        rp, col, val = blk.rows_global
        yg = orc.csr_spmv(rp.astype(np.int32), col.astype(np.int32), val, xg)
        ag = orc.csr_abs_spmv(rp.astype(np.int32), col.astype(np.int32), val, xg)
        assert np.all(np.abs(y_nat - yg) <= 1e-12 * ag), np.abs(y_nat - yg).max()
    else:
        # global reference: the whole stencil, natural order
        rp, col, val = mg.gen_stencil27_rows(grid[0], grid[1], grid[2] * world, 0, grid[2] * world)
        yg = orc.csr_spmv(rp.astype(np.int32), col.astype(np.int32), val, xg)
        ag = orc.csr_abs_spmv(rp.astype(np.int32), col.astype(np.int32), val, xg)
        assert np.all(np.abs(y_nat - yg[r0:r1]) <= 1e-12 * ag[r0:r1]), np.abs(y_nat - yg[r0:r1]).max()
    blk.free()
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
