#!/usr/bin/env python
"""Development tool: per-CTA timeline of one distributed product (EHYB_TRACE=1), summarised.

  EHYB_TRACE=1 torchrun --nproc-per-node N scripts/mg_trace.py [tag]     (N >= 2: ehyb_mg_*, peer-memory exchange)
  EHYB_TRACE=1 python scripts/mg_trace.py [tag]                          (one GPU: plain session)

Prints where the time of the last of 200 back-to-back products went on every rank and saves the
raw stamps to gpurun_out/trace_<tag>_rank<r>.npy (8 words per CTA, see ehyb_trace_read)."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("EHYB_MTMETIS_BIN", str(ROOT / "bin" / "ehyb_mtmetis"))
os.environ["EHYB_TRACE"] = "1"
GRID = (128, 128, 128)


def read_trace(lib, handle):
    n = C.c_int()
    lib.ehyb_trace_read(handle, None, C.byref(n))
    buf = np.zeros(n.value * 8, np.uint64)
    lib.ehyb_trace_read(handle, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), C.byref(n))
    return buf.reshape(-1, 8)


def summarise_persistent(tag, rank, tr, ms_per):
    """persistent kernel: one CTA per SM for the whole product (stamps: see ehyb_persistent_kernel)"""
    tr = tr[tr[:, 0] > 0]
    t = tr.astype(np.float64)
    T0 = t[:, 1].min()
    start, dep, push, staged, end = ((t[:, i] - T0) / 1e3 for i in range(5))
    out = ["[%s rank %d] persistent kernel, %d CTAs: %.2f us/product; span (first dep-wait passed -> last warp done) %.1f us"
           % (tag, rank, len(t), ms_per * 1e3, end.max())]
    out.append("  CTA start before its dep-wait: mean %.1f us (prologue overlapped with the previous product); dep-wait passed %.2f..%.2f"
               % ((dep - start).mean(), dep.min(), dep.max()))
    out.append("  first window staged %.2f..%.2f us after the dep-wait (mean %.2f)" % ((staged - dep).min(), (staged - dep).max(), (staged - dep).mean()))
    out.append("  CTA end: %.1f..%.1f (mean %.1f) -> idle tail mean %.1f us" % (end.min(), end.max(), end.mean(), (end.max() - end).mean()))
    pushed = tr[:, 2] > 0
    if pushed.any():
        out.append("  push: %d CTAs, done %.2f..%.2f us after their dep-wait (mean %.2f)"
                   % (pushed.sum(), (push - dep)[pushed].min(), (push - dep)[pushed].max(), (push - dep)[pushed].mean()))
    waited = tr[:, 6] > 0
    if waited.any():
        out.append("  halo wait: %d CTAs, longest wait per CTA %.2f..%.2f us (mean %.2f), ended at %.1f..%.1f us of the product"
                   % (waited.sum(), tr[waited, 6].min() / 1e3, tr[waited, 6].max() / 1e3, tr[waited, 6].mean() / 1e3,
                      ((t[:, 7] - T0) / 1e3)[waited].min(), ((t[:, 7] - T0) / 1e3)[waited].max()))
    print("\n".join(out), flush=True)


def summarise(tag, rank, tr, ms_per, halo_parts=None):
    if os.environ.get("EHYB_TRACE_PERSISTENT"):
        return summarise_persistent(tag, rank, tr, ms_per)
    t = tr.astype(np.float64)
    T0 = t[:, 1].min()               # first CTA past the wait for the previous product
    start, dep, push, staged, end = ((t[:, i] - T0) / 1e3 for i in range(5))
    sm = tr[:, 5].astype(int)
    first = np.zeros(len(t), bool)
    # first CTA on every SM = first wave
    for i in np.argsort(t[:, 0]):
        pass
    order = np.argsort(t[:, 0])
    seen = set()
    for i in order:
        if sm[i] not in seen:
            seen.add(sm[i]); first[i] = True
    out = ["[%s rank %d] %.2f us/product; kernel span (first dep-wait passed -> last warp done) %.1f us" % (tag, rank, ms_per * 1e3, end.max())]
    for name, sel in (("wave 1", first), ("later ", ~first)):
        if not sel.any():
            continue
        out.append("  %s: %3d CTAs  start %6.1f..%6.1f  staged-after-dep mean %5.2f max %5.2f  duration mean %5.1f max %5.1f  end %6.1f..%6.1f"
                   % (name, sel.sum(), dep[sel].min(), dep[sel].max(), (staged - dep)[sel].mean(), (staged - dep)[sel].max(),
                      (end - dep)[sel].mean(), (end - dep)[sel].max(), end[sel].min(), end[sel].max()))
    cr = tr[:, 7] > 0
    if cr.any():
        cache = (t[:, 7] - T0) / 1e3
        out.append("  remainder cache needed+ready (warp 0's first remainder chunk) %.2f..%.2f us after the dep-wait, mean %.2f"
                   % ((cache - dep)[cr].min(), (cache - dep)[cr].max(), (cache - dep)[cr].mean()))
    pushed = tr[:, 2] > 0
    if pushed.any():
        out.append("  push: %d CTAs, done %.2f..%.2f us after their dep-wait (last at %.2f us of the product)"
                   % (pushed.sum(), (push - dep)[pushed].min(), (push - dep)[pushed].max(), push[pushed].max()))
    if halo_parts is not None and len(halo_parts):
        hp = np.isin(tr[:, 6].astype(int), halo_parts)
        out.append("  halo partitions: %d CTAs, start %.1f..%.1f, staged-after-dep mean %.2f max %.2f, end max %.1f"
                   % (hp.sum(), dep[hp].min(), dep[hp].max(), (staged - dep)[hp].mean(), (staged - dep)[hp].max(), end[hp].max()))
    print("\n".join(out), flush=True)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "t"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    torch.cuda.set_device(local)
    from ehyb_spmv_gpu_b200 import _lib as L
    from ehyb_spmv_gpu_b200 import api
    lib = L.load()
    lib.ehyb_trace_read.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_int)]
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    if world == 1:
        from bench import build_matrix, stdout_to_stderr
        with stdout_to_stderr():
            m, lay, x, pl, _ = build_matrix(GRID)
        s = api.Session(lay, device=local)
        s.set_x(m.vector_reorder(x))
        ms = s.time_spmv(10, 200)
        ms = ms[0] if isinstance(ms, tuple) else ms
        tr = read_trace(lib, s.h)
        np.save(out_dir / ("trace_%s_rank0.npy" % tag), tr)
        if s.kernel_name() == "ehyb_persistent_kernel":
            os.environ["EHYB_TRACE_PERSISTENT"] = "1"
        summarise(tag, 0, tr, ms / 200)
        return
    import torch.distributed as dist
    from ehyb_spmv_gpu_b200 import multigpu as mg
    from bench import stdout_to_stderr
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    with stdout_to_stderr():
        blk, rowStarts = mg.setup_slab(rank, world, GRID, dist, "metis", "p2p")
        blk.create_session_p2p(local, dist)
    r0 = int(rowStarts[rank])
    x_perm = np.empty(blk.n)
    x_perm[blk.coo["reorderList"]] = mg.x_of_global(np.arange(r0, r0 + blk.n))
    blk.set_x(x_perm)
    dist.barrier()
    torch.cuda.synchronize()
    ms = blk.time_spmv(10, 200)
    torch.cuda.synchronize()
    tr = read_trace(lib, blk.handle)
    np.save(out_dir / ("trace_%s_rank%d.npy" % (tag, rank)), tr)
    lib.ehyb_session_kernel.restype = C.c_char_p
    if lib.ehyb_session_kernel(blk.handle) == b"ehyb_persistent_kernel":
        os.environ["EHYB_TRACE_PERSISTENT"] = "1"
    # partitions whose remainder cache holds halo columns
    v = api.LayoutView()
    lib.ehyb_layout_get(blk.layout, C.byref(v))
    parts = api._np(v.parts, v.nParts * 8, np.int32).reshape(v.nParts, 8)
    cc = api._np(v.cacheCols, v.cacheTotal, np.int32)
    halo_parts = [p for p in range(v.nParts) if parts[p, 5] > 0 and cc[parts[p, 4] + parts[p, 5] - 1] >= blk.n]
    for r in range(world):
        if r == rank:
            summarise(tag, rank, tr, ms / 200, np.array(halo_parts))
        dist.barrier()
    dist.barrier()
    blk.free()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
