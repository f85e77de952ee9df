"""Worker of tests/test_gpu_multi.py: one rank (one GPU) of a multi-GPU job.

Runs distributed products through the C ABI (ehyb_mg_*) with a DIFFERENT x every product - a
halo value read too early, too late or from the wrong buffer parity cannot pass - and checks
every y against the oracle's CSR product of the rank's block on [x_local | halo]."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("EHYB_MTMETIS_BIN", str(ROOT / "bin" / "ehyb_mtmetis"))


def _ndev():
    import torch
    return max(1, torch.cuda.device_count())


def setup_oneway(rank, world, dist, exchange, n_local=6000):
    """A block lower-bidiagonal random matrix: rank r references its own columns and columns of
    rank r-1 only.  The halo dependencies are ONE-WAY (rank 0 sends and never receives, the last
    rank receives and never sends): the flag protocol must still pair every sender with its
    receiver for the reuse of the halo buffers."""
    from ehyb_spmv_gpu_b200 import api
    from ehyb_spmv_gpu_b200 import multigpu as mg
    rng = np.random.default_rng(100 + rank)
    rowStarts = np.arange(world + 1, dtype=np.int64) * n_local
    r0 = int(rowStarts[rank])
    cols, vals, rowPtr = [], [], [0]
    for i in range(n_local):
        own = np.unique(np.concatenate([[i], rng.integers(max(0, i - 40), min(n_local, i + 40), 6)])) + r0
        ext = np.unique(rng.integers(0, n_local, 3)) + (r0 - n_local) if rank > 0 and i % 3 == 0 else np.zeros(0, np.int64)
        c = np.concatenate([ext, own]).astype(np.int64)
        cols.append(c); vals.append(rng.uniform(-1, 1, len(c))); rowPtr.append(rowPtr[-1] + len(c))
    rowPtr, cols, vals = np.array(rowPtr, np.int64), np.concatenate(cols), np.concatenate(vals)
    blk = mg.DistributedBlock(rank, world, rowStarts, rowPtr, cols, vals)
    blk.rows_global = (rowPtr, cols, vals)   # for reference products in the tests
    blk.exchange_lists(dist)
    assert (blk.nHalo > 0) == (rank > 0) and (int(blk.sendCount.sum()) > 0) == (rank < world - 1)
    try:
        dev = api.device_query(rank % max(1, _ndev()))
    except Exception:                        # the CPU (gloo) tier: nominal B200
        dev = api.device_info_b200()
    pl = api.plan(blk.n, dev)
    blk.finish(pl.nParts, pl.W, pl.ctasPerPart, None, exchange=exchange)
    return blk, rowStarts


def _same_reduced(blobs, world):
    """the reduced values (first 1 + i % 4 of row i) are bit-identical on every rank"""
    arrs = [np.frombuffer(b, np.float64).reshape(7, 4) for b in blobs]
    return all(np.array_equal(a[i, :1 + i % 4], arrs[0][i, :1 + i % 4]) for a in arrs for i in range(7))


def run_grid(rank, world, dist, exchange, grid, level1, products):
    """BASELINE.json config 5's pipeline at a small size: GLOBAL grid cut into bricks, level 1 by
    mt-metis on the brick graph (or scattered owners: every rank neighbours every other), streamed
    format build, products checked against the closed-form check vector of the stencil (oracle)."""
    import torch
    from ehyb_spmv_gpu_b200 import _lib as L
    from ehyb_spmv_gpu_b200 import multigpu as mg
    from oracle import oracle as O
    brick = (8, 8, 4)
    if level1 == "scatter":
        nb = int(np.prod([-(-g // b) for g, b in zip(grid, brick)]))
        dec = mg.GridDecomp(grid, brick, world, np.random.default_rng(7).integers(0, world, nb).astype(np.uint32))
        blk = mg.GridBlock(dec, rank, exchange=exchange, chunk_bricks=5)
        blk.exchange_lists(dist)
        assert int(np.count_nonzero(blk.recvCount)) == world - 1
    else:
        blk, dec = mg.setup_grid(rank, world, grid, brick, dist, "metis", exchange)
    if exchange == "p2p":
        blk.create_session_p2p(rank % _ndev(), dist)
        assert blk.launches_per_spmv() == 1 + (1 if blk.stats["nOverflow"] else 0)
    else:
        ids = [mg.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        blk.create_session(rank, ids[0])
    orc = O.Oracle()
    nat = blk.natural_ids()
    lib = blk.lib
    for k in range(products):           # host-synchronised products, a different x every time
        blk.set_x(orc.x_of_global(nat, k + 1.0, 0.01 * k))
        blk.spmv()
        y = blk.get_y()
        y_ref, absAx = orc.stencil27_rows_product(grid, nat, k + 1.0, 0.01 * k)
        bad = np.flatnonzero(~(np.abs(y - y_ref) <= 1e-12 * absAx))
        assert bad.size == 0, "rank %d product %d: %d rows outside the gate, first %s" % (rank, k, bad.size, bad[:5])
    # queued products without host synchronisation, every one with its own x (see main())
    K = 6
    xs = torch.zeros((K, blk.n + blk.nHalo), dtype=torch.float64, device="cuda")
    ys = torch.full((K, blk.n), float("nan"), dtype=torch.float64, device="cuda")
    for k in range(K):
        xs[k, :blk.n] = torch.from_numpy(orc.x_of_global(nat, 0.5 + k, -0.02 * k))
    torch.cuda.synchronize()
    dist.barrier()
    for k in range(K):
        L.check(lib, lib.ehyb_mg_spmv(blk.session, C.c_void_p(xs[k].data_ptr()), C.c_void_p(ys[k].data_ptr())), "ehyb_mg_spmv")
    L.check(lib, lib.ehyb_sync(blk.handle), "ehyb_sync")
    torch.cuda.synchronize()
    for k in range(K):
        y_ref, absAx = orc.stencil27_rows_product(grid, nat, 0.5 + k, -0.02 * k)
        bad = np.flatnonzero(~(np.abs(ys[k].cpu().numpy() - y_ref) <= 1e-12 * absAx))
        assert bad.size == 0, "rank %d queued product %d: %d rows outside the gate" % (rank, k, bad.size)
    # pipelined host-vector products (ehyb_mg_spmv_host_batch): y_i of x_i, every i
    xh = [orc.x_of_global(nat, 1.0 + 0.25 * i, 0.003 * i) for i in range(5)]
    yh = [np.full(blk.n, np.nan) for _ in range(5)]
    xp = (C.c_void_p * 5)(*[a.ctypes.data for a in xh])
    yp = (C.c_void_p * 5)(*[a.ctypes.data for a in yh])
    dist.barrier()
    L.check(lib, lib.ehyb_mg_spmv_host_batch(blk.session, xp, yp, 5), "ehyb_mg_spmv_host_batch")
    for i in range(5):
        y_ref, absAx = orc.stencil27_rows_product(grid, nat, 1.0 + 0.25 * i, 0.003 * i)
        assert np.all(np.abs(yh[i] - y_ref) <= 1e-12 * absAx), "rank %d host batch product %d outside the gate" % (rank, i)
    assert not blk.timed_out()
    # scalar all-reduce over peer memory (the dots of a distributed solver): 7 of them queued without host
    # synchronisation, every one with its own values; every rank must hold the SAME bits afterwards
    vals = torch.zeros((7, 4), dtype=torch.float64, device="cuda")
    for i in range(7):
        vals[i] = torch.tensor([(rank + 1) * (i + 1), 0.1 * (rank + 1) + i, 1e-3 * (rank - 0.5) * (i + 2), float(rank == i % world)], dtype=torch.float64)
    torch.cuda.synchronize()
    dist.barrier()
    for i in range(7):
        blk.allreduce_sum(vals[i].data_ptr(), 1 + i % 4)
    L.check(lib, lib.ehyb_sync(blk.handle), "ehyb_sync")
    torch.cuda.synchronize()
    got = vals.cpu().numpy()
    for i in range(7):
        cnt = 1 + i % 4
        mine = np.array([(rank + 1) * (i + 1), 0.1 * (rank + 1) + i, 1e-3 * (rank - 0.5) * (i + 2), float(rank == i % world)])
        want = np.zeros(4)
        for g in range(world):                      # rank order, as the kernel adds
            want += np.array([(g + 1) * (i + 1), 0.1 * (g + 1) + i, 1e-3 * (g - 0.5) * (i + 2), float(g == i % world)])
        if exchange == "p2p":   # the mailbox adds in rank order: these exact bits; NCCL picks its own order
            assert np.array_equal(got[i, :cnt], want[:cnt]), "rank %d all-reduce %d: %s != %s" % (rank, i, got[i, :cnt], want[:cnt])
        else:
            assert np.allclose(got[i, :cnt], want[:cnt], rtol=1e-14, atol=0), "rank %d all-reduce %d: %s != %s" % (rank, i, got[i, :cnt], want[:cnt])
        assert np.array_equal(got[i, cnt:], mine[cnt:]), "rank %d all-reduce %d touched values beyond its count" % (rank, i)
    everybody = [None] * world
    dist.all_gather_object(everybody, got.tobytes())
    assert _same_reduced(everybody, world), "the ranks hold different sums"
    assert not blk.timed_out()
    # distributed Jacobi-PCG: A x = b with b = A x_true from the closed form; all ranks stop in the same iteration
    x_true = orc.x_of_global(nat, 1.0, 0.0)
    b_loc, _ = orc.stencil27_rows_product(grid, nat, 1.0, 0.0)
    for jacobi in (True, False):
        xs_, info = blk.pcg_solve(b_loc, np.full(blk.n, 26.0) if jacobi else None, max_iters=3000, rtol=1e-10, check_every=4)
        infos = [None] * world
        dist.all_gather_object(infos, (info["iters"], info["converged"], info["rel_residual"], info["true_rel_residual"]))
        assert all(i == infos[0] for i in infos), "the ranks disagree about the solve: %s" % (infos,)
        assert info["converged"] and info["true_rel_residual"] <= 1e-8, info
        err2 = torch.tensor([float(np.sum((xs_ - x_true) ** 2)), float(np.sum(x_true ** 2))], dtype=torch.float64)
        dist.all_reduce(err2)
        assert float(err2[0]) <= (1e-6 ** 2) * float(err2[1]), "rank %d: PCG solution off: %s" % (rank, err2)
    assert not blk.timed_out()
    ms = blk.time_spmv(3, 50)
    kname = blk.kernel_name()
    dist.barrier()
    blk.free()
    dec.free()
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok (grid path, kernel %s, %d peers, %.1f us/product; distributed PCG %d iterations, %.1f us each)"
          % (rank, kname, int(np.count_nonzero(blk.recvCount)), ms / 50 * 1e3, info["iters"], info["ms"] * 1e3 / max(info["iters"], 1)))


def main():
    import torch
    import torch.distributed as dist
    from ehyb_spmv_gpu_b200 import _lib as L
    from ehyb_spmv_gpu_b200 import multigpu as mg
    from oracle import oracle as O

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    exchange, partition = sys.argv[1], sys.argv[2]
    grid = tuple(int(v) for v in sys.argv[3].split("x")) if "x" in sys.argv[3] else None
    products = int(sys.argv[4]) if len(sys.argv) > 4 else 7
    torch.cuda.set_device(rank % _ndev())   # (more ranks than GPUs: EHYB_MG_SHARE_DEVICE=1, ranks share a GPU)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if exchange == "p2p":
        assert mg.p2p_supported(rank % _ndev(), world), "no peer access between the GPUs of this box"
    if partition.startswith("grid"):
        return run_grid(rank, world, dist, exchange, grid, partition.split("-")[1], min(products, 4))
    if partition == "oneway":
        blk, rowStarts = setup_oneway(rank, world, dist, exchange)
    else:
        blk, rowStarts = mg.setup_slab(rank, world, grid, dist, partition, exchange)
    if exchange == "p2p":
        blk.create_session_p2p(rank % _ndev(), dist)
        assert blk.launches_per_spmv() == 1 + (1 if blk.stats["nOverflow"] else 0)
    else:
        ids = [mg.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        blk.create_session(rank, ids[0])
    orc = O.Oracle()
    r0 = int(rowStarts[rank])
    gl = np.arange(r0, r0 + blk.n)
    worst = 0.0
    for k in range(products):
        # x_k(global i) = base(i) * (k + 1) + 0.01 k: every product has its own halo values
        x_nat = mg.x_of_global(gl) * (k + 1) + 0.01 * k
        x_perm = np.empty(blk.n)
        x_perm[blk.coo["reorderList"]] = x_nat
        blk.set_x(x_perm)
        blk.spmv()
        y = blk.get_y()
        halo = mg.x_of_global(blk.haloGlobal) * (k + 1) + 0.01 * k
        x_ext = np.concatenate([x_perm, halo])
        y_ref = orc.csr_spmv(blk.coo["rowIdx"], blk.coo["J"], blk.coo["V"], x_ext)
        absAx = orc.csr_abs_spmv(blk.coo["rowIdx"], blk.coo["J"], blk.coo["V"], x_ext)
        bad = np.flatnonzero(~(np.abs(y - y_ref) <= 1e-12 * absAx))
        assert bad.size == 0, "rank %d product %d: %d rows outside the gate, first %s" % (rank, k, bad.size, bad[:5])
        worst = max(worst, float(np.abs(y - y_ref).max()))
        x_perm_last, x_ext_last = x_perm, x_ext
    assert not blk.timed_out(), "a neighbour did not deliver its halo in time"
    # K products queued back to back WITHOUT any host synchronisation in between, every one with
    # its own x_k (what a solver does): product k+1 starts early (programmatic dependent launch), a
    # sender may run ahead of its receiver, and the halo buffers alternate - a halo value of the
    # wrong product in any y_k fails the gate of x_k
    K = 6
    lib = blk.lib
    xs = torch.zeros((K, blk.n + blk.nHalo), dtype=torch.float64, device="cuda")
    ys = torch.full((K, blk.n), float("nan"), dtype=torch.float64, device="cuda")
    x_exts = []
    for k in range(K):
        x_nat = mg.x_of_global(gl) * (0.5 + k) - 0.02 * k
        x_perm = np.empty(blk.n)
        x_perm[blk.coo["reorderList"]] = x_nat
        xs[k, :blk.n] = torch.from_numpy(x_perm)
        x_exts.append(np.concatenate([x_perm, mg.x_of_global(blk.haloGlobal) * (0.5 + k) - 0.02 * k]))
    torch.cuda.synchronize()
    dist.barrier()
    for k in range(K):
        L.check(lib, lib.ehyb_mg_spmv(blk.session, C.c_void_p(xs[k].data_ptr()), C.c_void_p(ys[k].data_ptr())), "ehyb_mg_spmv")
    L.check(lib, lib.ehyb_sync(blk.handle), "ehyb_sync")
    torch.cuda.synchronize()
    for k in range(K):
        yk = ys[k].cpu().numpy()
        y_ref = orc.csr_spmv(blk.coo["rowIdx"], blk.coo["J"], blk.coo["V"], x_exts[k])
        absAx = orc.csr_abs_spmv(blk.coo["rowIdx"], blk.coo["J"], blk.coo["V"], x_exts[k])
        bad = np.flatnonzero(~(np.abs(yk - y_ref) <= 1e-12 * absAx))
        assert bad.size == 0, "rank %d queued product %d: %d rows outside the gate, first %s" % (rank, k, bad.size, bad[:5])
    # restore the last checked x / y of the loop above for the timed loop below
    blk.set_x(x_perm_last)
    blk.spmv()
    y = blk.get_y()
    y_ref = orc.csr_spmv(blk.coo["rowIdx"], blk.coo["J"], blk.coo["V"], x_ext_last)
    absAx = orc.csr_abs_spmv(blk.coo["rowIdx"], blk.coo["J"], blk.coo["V"], x_ext_last)
    # back-to-back products of the same x (the timed loop)
    ms = blk.time_spmv(3, 50)
    assert ms > 0
    y2 = blk.get_y()
    if blk.stats["nOverflow"] == 0:
        # every row is summed by one lane in a fixed order: bit-reproducible
        assert np.array_equal(y2, y), "back-to-back products of the same x changed y"
    else:
        # overflow entries are added with atomics (order varies): inside the gate of the last x
        assert np.all(np.abs(y2 - y_ref) <= 1e-12 * absAx), "back-to-back products left the gate"
    dist.barrier()
    blk.free()
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok (max abs err %.3g, %.1f us/product)" % (rank, worst, ms / 50 * 1e3))


if __name__ == "__main__":
    main()
