#!/usr/bin/env python
"""bench.py -- fp64 EHYB SpMV throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step is one y = A x product.
  N = 1  BASELINE.json configs[1]: 3-D 27-point stencil 128^3 (n 2 097 152, nnz 55 742 968, fp64) on one
         GPU - the configuration the metric is quoted on.
  N > 1  BASELINE.json configs[4]: 3-D 27-point stencil 512^3 (n 134 217 728, nnz 3 609 741 304) sharded
         over the N GPUs (strong scaling: the global problem is fixed).  The grid is cut into 16^3
         bricks; level 1 = the pinned mt-metis k = N partition of the weighted brick graph (metis row
         blocks -> GPUs), level 2 = one EHYB partition per brick; every rank streams its block into the
         tuned layout without ever holding a COO (csrc/host/grid.c).  The x halo is exchanged every
         product INSIDE the persistent kernel (stores into the neighbours' memory over NVLink, CUDA
         IPC; EHYB_MG_EXCHANGE=nccl selects the NCCL send/recv baseline).
Inputs are resident in HBM when the timed region starts; the matrix data per GPU (>= 580 MB) is
larger than L2 (126 MB).  EHYB_BENCH_GRID=NXxNYxNZ sets another global grid (also at N = 1: the same
brick pipeline on one GPU), EHYB_BENCH_BRICK the brick, EHYB_BENCH_LEVEL1=runs contiguous runs of
bricks instead of mt-metis; EHYB_BENCH_MG=slab runs round 1's weak-scaling z-slab case (128^3 per GPU).

The JSON line carries: value (GFLOP/s = 2 nnz / t, all ranks), roofline (algorithmic bytes of
the dominant kernel / its CUDA-event duration, against MEASURED_PEAKS.json; traffic = DRAM bytes
per launch from the committed ncu capture), e2e (the same metric through the host-buffer entry
point, H2D of x and D2H of y inside the timed region), cpu_baseline (the oracle's CSR product on
the host cores; rank 0, N=1 only), comparisons (after the timed region, same GPU: cuSPARSE CSR, the
reference's own kernel.cu recompiled for sm_100a, and BASELINE.json configs[0] - 5-point 1024^2,
L2-resident - with a warm and with a flushed L2), clocks.  `--impl reference` times the reference's CPU path (CSR over its own arrays, all
host threads; at N > 1 on a bounded sample of configs[4]: a 512 x 512 x 8 slab of the stencil).  The oracle is used only as checker and CPU baseline, never on the product path.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("EHYB_MTMETIS_BIN", str(ROOT / "bin" / "ehyb_mtmetis"))

GRID = (128, 128, 128)  # BASELINE.json configs[1]
WORKLOAD = "3D 27-point stencil 128^3 (n 2097152, nnz 55742968) fp64, EHYB, single B200 (BASELINE.json configs[1])"
GRID5 = (512, 512, 512)  # BASELINE.json configs[4]
BRICK = tuple(int(v) for v in os.environ.get("EHYB_BENCH_BRICK", "16x16x16").lower().split("x"))
SLAB = os.environ.get("EHYB_BENCH_MG") == "slab"  # round 1's weak-scaling case: 128^3 per GPU, z-slabs
if os.environ.get("EHYB_BENCH_GRID"):
    GRID = GRID5 = tuple(int(v) for v in os.environ["EHYB_BENCH_GRID"].lower().split("x"))
STRONG = os.environ.get("EHYB_BENCH_SCALING") == "strong"


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = get_reasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


class stdout_to_stderr:
    """The C library prints the reference's log lines on stdout; keep stdout for the JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def build_matrix(grid, x=None):
    """generate -> matrixCOO -> plan -> mt-metis -> reorder -> tuned layout (all product code)."""
    from ehyb_spmv_gpu_b200 import api
    t0 = time.time()
    n, li, lj, lv = api.gen_lower(api.GEN_STENCIL27, *grid)
    if x is None:
        x = api.x_reference(n)
    m = api.CooMatrix.from_lower(n, li, lj, lv, x)
    del li, lj, lv
    try:
        dev = api.device_query(int(os.environ.get("LOCAL_RANK", "0")))
    except Exception:
        dev = api.device_info_b200()
    pl = api.plan(n, dev, kernel=api.KERNEL_PERSISTENT)  # single GPU: partitions sized for the persistent kernel
    # partition stage: deterministic and parallel (hierpart.c: 8 single-threaded mt-metis processes at once)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    api.L.load().ehyb_set_partition_pieces(int(os.environ.get("EHYB_PARTITION_PIECES", min(8, cores))))
    m.set_plan(pl.nParts, pl.W, pl.ctasPerPart)
    m.reorder()
    lay = api.Layout(m, er_fill=float(os.environ.get("EHYB_ER_FILL", "-1")))
    return m, lay, x, pl, time.time() - t0


def config1_l2_row(device):
    """BASELINE.json configs[0] (5-point 1024^2: 20 MB of matrix data, L2-resident) with a warm and with
    a cold L2 (SURVEY.md 8d "Timing"); reported next to the headline, outside its timed region."""
    from ehyb_spmv_gpu_b200 import api
    try:
        n, li, lj, lv = api.gen_lower(api.GEN_LAPLACE2D, 1024, 1024, 1)
        x = api.x_reference(n)
        m = api.CooMatrix.from_lower(n, li, lj, lv, x)
        pl, kern = api.plan_auto(n, m.nnz, api.device_query(device))   # L2-resident: one partition per SM, staged kernel
        m.set_plan(pl.nParts, pl.W, pl.ctasPerPart)
        m.reorder()
        lay = api.Layout(m, er_fill=0.0)                              # a second launch for a few entries costs more than their padding
        st = lay.stats()
        s = api.Session(lay, device=device, kernel=kern)
        xr = m.vector_reorder(x)
        s.set_x(xr)
        iters = 200
        warm = s.time_spmv(10, iters) / iters * 1e3
        cold = s.time_spmv_flushed(3, 50) / 50 * 1e3
        ok = bool(np.all(np.abs(m.vector_recover(s.get_y()) - m.y_golden) <= 1e-12 * np.maximum(np.abs(m.y_golden), 1.0)))
        row = {"workload": "2D 5-point Laplacian 1024^2 (n %d, nnz %d), BASELINE.json configs[0]" % (n, st["nnz"]),
               "kernel": s.kernel_name(), "partitions": st["nParts"], "window": st["W"], "format_bytes": st["formatBytes"],
               "us_per_product_l2_resident": round(warm, 2), "GBs_algorithmic_l2_resident": round(st["algBytes"] / (warm * 1e3), 1),
               "us_per_product_l2_flushed": round(cold, 2), "GBs_algorithmic_l2_flushed": round(st["algBytes"] / (cold * 1e3), 1),
               "what": "l2_resident: %d back-to-back products (programmatic dependent launch); l2_flushed: 512 MB of scratch "
                       "overwritten before every product, each product timed by its own event pair (plain launches)" % iters,
               "result_inside_gate": ok}
        s.free(); lay.free(); m.free()
        return row
    except Exception as e:  # a comparison row must not take the bench line down
        return {"error": str(e)[:200]}


def gpu_comparisons(m, xr, y_ref, absAx):
    """Other GPU implementations of the same product on the same device, outside the timed region
    (reported rows): cuSPARSE generic-API CSR SpMV through spmvGeneric's library
    (libehyb_cusparse.so, the reference's comparison path made to work).  Empty if not built."""
    out = {}
    path = ROOT / "ehyb_spmv_gpu_b200" / "lib" / "libehyb_cusparse.so"
    if not path.exists():
        return out
    try:
        lib = C.CDLL(str(path))
        lib.ehyb_cusparse_last_error.restype = C.c_char_p
        nnz = m.nnz
        for alg in (1, 2):
            us = C.c_float()
            y = np.empty(m.n)
            rc = lib.ehyb_cusparse_spmv(C.byref(m.c), xr.ctypes.data_as(C.POINTER(C.c_double)),
                                        y.ctypes.data_as(C.POINTER(C.c_double)), 5, 50, alg, C.byref(us))
            if rc != 0:
                out["cusparse_csr_alg%d" % alg] = {"error": lib.ehyb_cusparse_last_error().decode()}
                continue
            out["cusparse_csr_alg%d" % alg] = {
                "us_per_product": round(us.value, 2), "GFLOP/s": round(2.0 * nnz / (us.value * 1e3), 1),
                "rows_outside_1e-12_gate": int(np.count_nonzero(~(np.abs(y - y_ref) <= 1e-12 * absAx))),
                "what": "cusparseSpMV CSR fp64 (CUSPARSE_SPMV_CSR_ALG%d, preprocessed) on the same permuted matrix, same GPU" % alg}
    except Exception as e:  # a comparison row must never take the bench down
        out["error"] = repr(e)
    return out


def reference_gpu_row(iters=100):
    """The reference's OWN device path on this GPU (reported row): reference kernel.cu +
    convert.c, unmodified, compiled in place for sm_100a (oracle/_ref/ref_gpu_bench, built by
    `make -C oracle refgpu`), at the reference's own partition parameters (82-SM heuristic), timed with
    CUDA events as shipped (the remainder phase only runs in the first launch) and with the
    remainder counter reset before every launch.  Outside the timed region; absent if not built."""
    exe = ROOT / "oracle" / "_ref" / "ref_gpu_bench"
    if not exe.exists():
        return {"unavailable": "oracle/_ref/ref_gpu_bench not built (needs /root/reference at build time)"}
    import subprocess
    import tempfile
    from oracle import oracle as O
    try:
        orc = O.Oracle()
        n, li, lj, lv = O.gen_stencil27_lower(*GRID)
        x = orc.x_reference(n)
        m = orc.read_sym(n, li, lj, lv, x)
        P, W, kpp = orc.heuristic_ref(n, True)
        xadj, adj = orc.graph(m)
        part = O.mtmetis_partition(xadj, adj, P, nthreads=1)
        r = orc.reorder(m, P, W, part)
        xr = orc.vector_reorder(x, r["reorderList"])
        with tempfile.TemporaryDirectory() as d:
            fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
            with open(fin, "wb") as f:
                f.write(np.array([n, m["nnz"], P, W, max(kpp, 1)], np.int32).tobytes())
                for k in ("I", "J"):
                    f.write(r[k].tobytes())
                f.write(r["V"].tobytes())
                for k in ("rowIdx", "numInRow", "numInRow2", "partBoundary"):
                    f.write(r[k].astype(np.int32).tobytes())
                f.write(xr.tobytes())
            res = subprocess.run([str(exe), fin, fout, str(iters)], capture_output=True, text=True, timeout=300)
            line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
            if res.returncode or not line:
                return {"error": (res.stdout + res.stderr)[-400:]}
            row = json.loads(line[-1])["reference_gpu"]
            yy = np.fromfile(fout, dtype=np.float64)
        y_ref = orc.csr_spmv(r["rowIdx"], r["J"], r["V"], xr)
        absAx = orc.csr_abs_spmv(r["rowIdx"], r["J"], r["V"], xr)
        row["rows_outside_1e-12_gate_as_shipped"] = int(np.count_nonzero(~(np.abs(yy[:n] - y_ref) <= 1e-12 * absAx)))
        row["rows_outside_1e-12_gate_repaired"] = int(np.count_nonzero(~(np.abs(yy[n:] - y_ref) <= 1e-12 * absAx)))
        row["partitions"], row["window"] = P, W
        row["what"] = ("reference kernel.cu:110-195 + convert.c, unmodified, nvcc -arch=sm_100a, same GPU, same matrix, the reference's "
                       "own partition parameters; 'as_shipped' = remainder work only in launch 1 (B-1), 'repaired' = every launch complete")
        return row
    except Exception as e:  # a comparison row must never take the bench down
        return {"error": repr(e)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ehyb_spmv_gpu_b200 import _lib, api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N > 1 must be launched with torchrun (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        with stdout_to_stderr():  # NCCL prints its version on stdout when the first communicator is made
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.all_reduce(torch.zeros(1, device="cuda"))
            torch.cuda.synchronize()
    lib = _lib.load()

    if world > 1 and SLAB:
        grid = GRID
        if STRONG:
            if GRID[2] % world:
                raise SystemExit("bench.py: strong scaling needs NZ divisible by the number of GPUs")
            grid = (GRID[0], GRID[1], GRID[2] // world)
        wl = "3D 27-point stencil %dx%dx%d fp64, EHYB, %s" % (GRID + ("global grid cut into z-slabs" if STRONG else "per GPU (z-slabs)",))
        return run_ours_multi(args, rank, world, local, grid, wl, "strong" if STRONG else "weak")
    if world > 1 or os.environ.get("EHYB_BENCH_GRID"):
        return run_ours_grid(args, rank, world, local)

    with stdout_to_stderr():
        m, lay, x, pl, t_prep = build_matrix(GRID)
    st = lay.stats()
    print("layout: P=%d W=%d slices=%d nnz_ell=%d in-slice remainder=%d overflow=%d" % (st["nParts"], st["W"], st["nSlices"], st["nnzEll"],
                                                                                      st["nnzRemInSlice"], st["nnzOverflow"]), file=sys.stderr, flush=True)
    s = api.Session(lay, device=local)
    xr = m.vector_reorder(x)
    s.set_x(xr)
    launches = s.launches_per_spmv()

    # ---- device-resident timed region: CUDA events on the session stream inside libehyb ----
    sampler = ClockSampler(local)
    sampler.start()
    torch.cuda.synchronize()
    ms, kms = s.time_spmv(args.warmup, args.steps, kernel_only=True)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_per_step = ms / args.steps
    gflops = 2.0 * st["nnz"] / (ms_per_step * 1e6)
    kernel_us = kms / args.steps * 1e3

    # quick parity check of what was timed (not in the timed region)
    y = m.vector_recover(s.get_y())
    err = float(np.abs(y - m.y_golden).max())

    # ---- end to end: host buffers, H2D of x and D2H of y every step, through the C ABI ----
    n = m.n
    pin = []
    def pinned(count):
        p = C.c_void_p()
        _lib.check(lib, lib.ehyb_host_alloc_pinned(C.c_size_t(count * 8), C.byref(p)), "ehyb_host_alloc_pinned")
        pin.append(p)
        return np.ctypeslib.as_array((C.c_double * count).from_address(p.value))
    nbuf = 4
    xs = [pinned(n) for _ in range(nbuf)]
    ys = [pinned(n) for _ in range(nbuf)]
    for b in xs:
        b[:] = xr
    s.spmv_host_batch([xs[i % nbuf] for i in range(max(args.warmup, 3))], [ys[i % nbuf] for i in range(max(args.warmup, 3))])
    t0 = time.perf_counter()
    s.spmv_host_batch([xs[i % nbuf] for i in range(args.steps)], [ys[i % nbuf] for i in range(args.steps)])
    t_e2e = time.perf_counter() - t0
    e2e_gflops = 2.0 * st["nnz"] * args.steps / t_e2e / 1e9
    y_again = s.spmv_host(np.asarray(xs[0]))
    if st["nnzOverflow"] == 0:
        # every row is summed by one lane in a fixed order: the pipelined path must reproduce it bit for bit
        assert np.array_equal(np.asarray(ys[0]), y_again)
    else:
        # overflow entries are added with atomics: equal up to summation order
        assert np.allclose(np.asarray(ys[0]), y_again, rtol=0, atol=1e-12 * float(np.abs(y_again).max() + 1.0))

    # ---- CPU baseline: the oracle's CSR product on the host cores (bounded sample) ----
    from oracle import oracle as O
    orc = O.Oracle()
    a = m.arrays()
    cpu_iters = 20
    sec, y_cpu = orc.csr_spmv_timed(a["rowIdx"], a["J"], a["V"], xr, 2, cpu_iters)
    cpu_gflops = 2.0 * st["nnz"] * cpu_iters / sec / 1e9
    absAx = orc.csr_abs_spmv(a["rowIdx"], a["J"], a["V"], xr)
    gate_fail = int(np.count_nonzero(~(np.abs(s.get_y() - y_cpu) <= 1e-12 * absAx)))

    comparisons = gpu_comparisons(m, xr, y_cpu, absAx)
    if GRID == (128, 128, 128) and os.environ.get("EHYB_BENCH_REFERENCE_GPU", "1") != "0":
        comparisons["reference_kernel_sm100a"] = reference_gpu_row()

    if GRID == (128, 128, 128) and os.environ.get("EHYB_BENCH_CONFIG1", "1") != "0":
        with stdout_to_stderr():
            comparisons["config1_l2"] = config1_l2_row(local)

    peaks, peak_src = measured_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # One product = `launches` kernels.  With a single launch per product the kernel's average
    # duration over the timed region IS the step time (consecutive launches overlap their
    # prologue with the predecessor's tail through programmatic dependent launch, so bracketing
    # every launch with its own events - kernel_us_isolated - serialises them and reads higher).
    kernel_us_timed = ms_per_step * 1e3 if launches == 1 else kernel_us
    achieved = st["algBytes"] / (kernel_us_timed * 1e3)  # bytes / ns = GB/s
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    out = {
        "metric": "fp64 SpMV GFLOP/s (2*nnz/t), EHYB format", "value": round(gflops, 2), "unit": "GFLOP/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 6),
        "higher_is_better": True, "scaling": "strong" if STRONG else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "grid": list(GRID), "n": n, "nnz": st["nnz"],
                   "partitions": st["nParts"], "window": st["W"], "ctas_per_partition": st["ctasPerPart"],
                   "slices": st["nSlices"], "nnz_ell": st["nnzEll"], "nnz_remainder_in_slice": st["nnzRemInSlice"],
                   "nnz_overflow": st["nnzOverflow"], "format_bytes": st["formatBytes"],
                   "l2": "matrix data (%.0f MB) larger than L2 (126 MB), no flush" % (st["formatBytes"] / 1e6),
                   "host_prep_s": round(t_prep, 1)},
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                     "kernel": s.kernel_name(), "kernel_us": round(kernel_us_timed, 3),
                     "kernel_us_isolated": round(kernel_us, 3),
                     "frac_isolated": round(st["algBytes"] / (kernel_us * 1e3) / peak, 4),
                     "note": "achieved = algorithmic bytes / average launch duration over the timed chain, where consecutive "
                             "launches overlap (programmatic dependent launch) and x / y stay in L2, so it can touch the measured "
                             "COPY peak; frac_isolated times every launch by its own event pair (serialised, no overlap)",
                     "algorithmic_bytes_per_launch": st["algBytes"],
                     "frac_of_nominal_8TBs": round(achieved / 8000.0, 4),
                     "whole_step_GBs": round(st["algBytes"] / (ms_per_step * 1e6), 1)},
        "e2e": {"value": round(e2e_gflops, 2), "unit": "GFLOP/s", "h2d_bytes_per_step": 8 * n,
                "d2h_bytes_per_step": 8 * n, "api": "ehyb_spmv_host_batch (pinned host x/y, copies pipelined)"},
        "cpu_baseline": {"value": round(cpu_gflops, 2), "unit": "GFLOP/s", "cores": orc.num_threads(), "kind": "port",
                         "sample": "%d CSR products of the same permuted matrix (oracle orc_csr_spmv, OpenMP)" % cpu_iters},
        "gpu_launches": launches * args.steps,
        "clocks": clocks,
        "parity": {"max_abs_err_vs_golden": err, "rows_outside_1e-12_gate": gate_fail},
        "comparisons": comparisons,
    }
    for p in pin:
        lib.ehyb_host_free_pinned(p)
    s.free(); lay.free(); m.free()
    print(json.dumps(out), flush=True)


EXCHANGE_TEXT = {
    "p2p": "inside the main kernel: x entries stored into the neighbours' halo buffers over NVLink (CUDA IPC peer "
           "memory, epoch flags), halo columns served from the shared-memory remainder cache; one launch per product",
    "nccl": "pack kernel + grouped ncclSend/ncclRecv per product, overlapped with the main kernel; halo entries in the "
            "overflow kernel",
}


def run_ours_multi(args, rank, world, local, grid, workload, scaling="weak"):
    """N > 1 (under torchrun, process group already initialised): one rank per GPU, distributed
    products through ehyb_mg_*; the oracle is used only for the parity check of one product."""
    import torch
    import torch.distributed as dist
    from ehyb_spmv_gpu_b200 import _lib as L
    from ehyb_spmv_gpu_b200._lib import check
    from ehyb_spmv_gpu_b200.multigpu import p2p_supported, setup_slab, unique_id, x_of_global

    t0 = time.time()
    exchange = os.environ.get("EHYB_MG_EXCHANGE", "p2p")
    if exchange == "p2p":
        # every rank must be able to map its neighbours' memory, else all fall back to NCCL
        ok = torch.tensor([1 if p2p_supported(local, world) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            raise SystemExit("bench.py: the GPUs of this box have no peer access; set EHYB_MG_EXCHANGE=nccl")
    with stdout_to_stderr():
        blk, rowStarts = setup_slab(rank, world, grid, dist, os.environ.get("EHYB_MG_PARTITION", "metis"), exchange)
        if exchange == "p2p":
            blk.create_session_p2p(local, dist)
        else:
            ids = [unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            blk.create_session(local, ids[0])
    t_prep = time.time() - t0
    r0 = int(rowStarts[rank])
    x_nat = x_of_global(np.arange(r0, r0 + blk.n))
    x_perm = np.empty(blk.n)
    x_perm[blk.coo["reorderList"]] = x_nat
    blk.set_x(x_perm)

    # parity of one distributed product: CPU CSR of the permuted local block on [x_local | halo]
    blk.spmv()
    y = blk.get_y()
    if blk.timed_out():
        raise SystemExit("bench.py: rank %d: a neighbour did not deliver its halo" % rank)
    from oracle import oracle as O
    orc = O.Oracle()
    x_ext = np.concatenate([x_perm, x_of_global(blk.haloGlobal)])
    y_ref = orc.csr_spmv(blk.coo["rowIdx"], blk.coo["J"], blk.coo["V"], x_ext)
    absAx = orc.csr_abs_spmv(blk.coo["rowIdx"], blk.coo["J"], blk.coo["V"], x_ext)
    gate_fail = int(np.count_nonzero(~(np.abs(y - y_ref) <= 1e-12 * absAx)))

    sampler = ClockSampler(local)
    sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    ms = blk.time_spmv(args.warmup, args.steps)
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    tot = torch.tensor([blk.stats["nnz"], blk.stats["algBytes"], gate_fail, blk.nHalo], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    nnz_all, alg_all, gate_all, halo_all = (float(v) for v in tot.tolist())

    # end to end: host x -> device, distributed product, y -> host, every step
    lib = blk.lib
    xe = np.zeros(blk.n + blk.nHalo); xe[:blk.n] = x_perm
    yh = np.empty(blk.n)
    xd = C.c_void_p(); yd = C.c_void_p()
    lib.ehyb_session_vectors(blk.handle, C.byref(xd), C.byref(yd))
    dist.barrier()
    te = time.perf_counter()
    for _ in range(args.steps):
        check(lib, lib.ehyb_set_x(blk.handle, xe.ctypes.data_as(L.c_dbl_p)), "ehyb_set_x")
        check(lib, lib.ehyb_mg_spmv(blk.session, xd, yd), "ehyb_mg_spmv")
        check(lib, lib.ehyb_get_y(blk.handle, yh.ctypes.data_as(L.c_dbl_p)), "ehyb_get_y")
    dist.barrier()
    te = time.perf_counter() - te

    ms_per_step = ms_max / args.steps
    if rank == 0:
        peaks, peak_src = measured_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = blk.stats["algBytes"] / (ms_per_step * 1e6)
        out = {
            "metric": "fp64 SpMV GFLOP/s (2*nnz/t), EHYB format", "value": round(2.0 * nnz_all / (ms_per_step * 1e6), 2),
            "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 6), "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "global_grid": [grid[0], grid[1], grid[2] * world],
                       "decomposition": "z-slabs, one per GPU; level-2 partition per GPU: " + os.environ.get("EHYB_MG_PARTITION", "metis"),
                       "n_per_gpu": blk.n, "nnz_total": int(nnz_all), "halo_x_entries_total": int(halo_all),
                       "partitions_per_gpu": blk.stats["nParts"], "window": blk.stats["W"],
                       "exchange": EXCHANGE_TEXT[exchange], "nnz_overflow_rank0": blk.stats["nOverflow"],
                       "remainder_cache_max": blk.stats["cacheMax"],
                       "l2": "matrix data per GPU larger than L2, no flush", "host_prep_s": round(t_prep, 1)},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                         "kernel": "ehyb_staged_kernel",
                         "note": "rank 0 algorithmic bytes / whole-step time (the step is the main kernel with the "
                                 "exchange inside it)" if exchange == "p2p" else
                                 "rank 0 algorithmic bytes / whole-step time (main kernel + exchange + overflow)"},
            "e2e": {"value": round(2.0 * nnz_all * args.steps / te / 1e9, 2), "unit": "GFLOP/s",
                    "h2d_bytes_per_step": 8 * (blk.n + blk.nHalo), "d2h_bytes_per_step": 8 * blk.n,
                    "api": "ehyb_set_x + ehyb_mg_spmv + ehyb_get_y per step, per rank"},
            "gpu_launches": args.steps * blk.launches_per_spmv(),
            "clocks": clocks,
            "parity": {"rows_outside_1e-12_gate_all_ranks": int(gate_all)},
        }
        print(json.dumps(out), flush=True)
    blk.free()
    dist.barrier()
    dist.destroy_process_group()


class _NoDist:
    """world == 1: the collectives of the set-up degenerate"""

    @staticmethod
    def all_gather_object(out, obj):
        out[0] = obj

    @staticmethod
    def barrier():
        pass

    @staticmethod
    def broadcast_object_list(box, src=0):
        pass


def run_ours_grid(args, rank, world, local):
    """BASELINE.json configs[4]: the 27-point stencil on GRID5 sharded over `world` GPUs by mt-metis
    blocks of bricks, streamed format build, halo exchange inside the persistent kernel.  The oracle
    (closed-form check vector) is used only for the parity check of one product."""
    import torch
    import torch.distributed as tdist
    from ehyb_spmv_gpu_b200 import _lib as L
    from ehyb_spmv_gpu_b200._lib import check
    from ehyb_spmv_gpu_b200.multigpu import p2p_supported, setup_grid, unique_id, x_of_global

    dist = tdist if world > 1 else _NoDist
    lib = L.load()
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    threads = max(1, cores // world)   # torchrun exports OMP_NUM_THREADS=1: the format build is OpenMP code
    lib.ehyb_set_host_threads(threads)
    t0 = time.time()
    exchange = os.environ.get("EHYB_MG_EXCHANGE", "p2p")
    if exchange == "p2p" and world > 1:
        ok = torch.tensor([1 if p2p_supported(local, world) else 0], device="cuda")
        tdist.all_reduce(ok, op=tdist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            raise SystemExit("bench.py: the GPUs of this box have no peer access; set EHYB_MG_EXCHANGE=nccl")
    level1 = os.environ.get("EHYB_BENCH_LEVEL1", "metis")
    with stdout_to_stderr():
        blk, dec = setup_grid(rank, world, GRID5, BRICK, dist, level1, exchange)
        t_build = time.time() - t0
        if exchange == "p2p":
            blk.create_session_p2p(local, dist)
        else:
            ids = [unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            blk.create_session(local, ids[0])
    t_prep = time.time() - t0
    nat = blk.natural_ids()
    x_perm = x_of_global(nat)
    blk.set_x(x_perm)

    # parity of one distributed product: the closed-form check vector of the stencil (oracle), every row
    blk.spmv()
    y = blk.get_y()
    if blk.timed_out():
        raise SystemExit("bench.py: rank %d: a neighbour did not deliver its halo" % rank)
    from oracle import oracle as O
    orc = O.Oracle()
    y_ref, absAx = orc.stencil27_rows_product(GRID5, nat)
    gate_fail = int(np.count_nonzero(~(np.abs(y - y_ref) <= 1e-12 * absAx)))
    max_err = float(np.abs(y - y_ref).max())
    del y_ref, absAx, nat

    sampler = ClockSampler(local)
    sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    ms = blk.time_spmv(args.warmup, args.steps)
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop()
    peers = int(np.count_nonzero(blk.recvCount))
    stats = torch.tensor([ms, blk.stats["algBytes"], peers, -peers, blk.nHalo, t_prep], dtype=torch.float64, device="cuda")
    tot = torch.tensor([blk.stats["nnz"], blk.stats["algBytes"], gate_fail, blk.nHalo, blk.stats["nnzOverflow"]], dtype=torch.float64, device="cuda")
    if world > 1:
        tdist.all_reduce(stats, op=tdist.ReduceOp.MAX)
        tdist.all_reduce(tot, op=tdist.ReduceOp.SUM)
    ms_max, alg_max, peers_max, neg_peers_min, halo_max, prep_max = (float(v) for v in stats.tolist())
    nnz_all, alg_all, gate_all, halo_all, ovf_all = (float(v) for v in tot.tolist())

    # end to end: host x -> device, distributed product, y -> host, every step, pipelined (pinned buffers)
    pin = []
    def pinned(count):
        p = C.c_void_p()
        check(lib, lib.ehyb_host_alloc_pinned(C.c_size_t(count * 8), C.byref(p)), "ehyb_host_alloc_pinned")
        pin.append(p)
        return np.ctypeslib.as_array((C.c_double * count).from_address(p.value))
    nbuf = 2
    xs = [pinned(blk.n) for _ in range(nbuf)]
    ys = [pinned(blk.n) for _ in range(nbuf)]
    for b in xs:
        b[:] = x_perm
    def host_batch(count):
        xp = (C.c_void_p * count)(*[xs[i % nbuf].ctypes.data for i in range(count)])
        yp = (C.c_void_p * count)(*[ys[i % nbuf].ctypes.data for i in range(count)])
        check(lib, lib.ehyb_mg_spmv_host_batch(blk.session, xp, yp, count), "ehyb_mg_spmv_host_batch")
    host_batch(3)
    e2e_steps = min(args.steps, 50)
    dist.barrier()
    te = time.perf_counter()
    host_batch(e2e_steps)
    dist.barrier()
    te = time.perf_counter() - te
    e2e_ok = bool(np.array_equal(np.asarray(ys[(e2e_steps - 1) % nbuf]), y)) if blk.stats["nOverflow"] == 0 else True
    tt = torch.tensor([te], dtype=torch.float64, device="cuda")
    if world > 1:
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
    te = float(tt.item())

    ms_per_step = ms_max / args.steps
    if rank == 0:
        peaks, peak_src = measured_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg_max / (ms_per_step * 1e6)
        n_all = int(np.prod(GRID5))
        is5 = GRID5 == (512, 512, 512)
        out = {
            "metric": "fp64 SpMV GFLOP/s (2*nnz/t), EHYB format", "value": round(2.0 * nnz_all / (ms_per_step * 1e6), 2),
            "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 6), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3D 27-point stencil %dx%dx%d (n %d, nnz %d) fp64, EHYB, sharded over %d B200%s"
                                   % (GRID5 + (n_all, int(nnz_all), world, " (BASELINE.json configs[4])" if is5 else "")),
                       "global_grid": list(GRID5), "brick": list(BRICK),
                       "decomposition": ("level 1: mt-metis k=%d partition of the weighted brick graph (bricks -> GPUs); " % world if level1 == "metis" and world > 1
                                         else "level 1: contiguous runs of bricks; ") + "level 2: one EHYB partition per brick",
                       "n_per_gpu_rank0": blk.n, "nnz_total": int(nnz_all), "halo_x_entries_total": int(halo_all),
                       "halo_x_entries_max_per_gpu": int(halo_max), "peers_per_gpu_min": int(-neg_peers_min), "peers_per_gpu_max": int(peers_max),
                       "partitions_rank0": blk.stats["nParts"], "window": blk.stats["W"],
                       "exchange": EXCHANGE_TEXT[exchange] if world > 1 else "none (one GPU)", "nnz_overflow_total": int(ovf_all),
                       "remainder_cache_max": blk.stats["cacheMax"],
                       "l2": "matrix data per GPU (%.1f GB on rank 0) larger than L2 (126 MB), no flush" % (blk.stats["formatBytes"] / 1e9),
                       "host_prep_s": round(prep_max, 1), "host_format_build_s_rank0": round(t_build, 1), "host_threads_per_rank": threads},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                         "kernel": blk.kernel_name(), "frac_of_nominal_8TBs": round(achieved / 8000.0, 4),
                         "algorithmic_bytes_per_launch": int(alg_max),
                         "note": "largest per-GPU algorithmic bytes / whole-step time (max over ranks); the step is one launch of "
                                 "the main kernel with the exchange inside it" if exchange == "p2p" else
                                 "largest per-GPU algorithmic bytes / whole-step time (main kernel + exchange + overflow)"},
            "e2e": {"value": round(2.0 * nnz_all * e2e_steps / te / 1e9, 2), "unit": "GFLOP/s",
                    "h2d_bytes_per_step": 8 * n_all, "d2h_bytes_per_step": 8 * n_all, "steps": e2e_steps,
                    "api": "ehyb_mg_spmv_host_batch per rank (pinned host x/y of the rank's rows, copies pipelined with the products)",
                    "bit_identical_to_device_resident_product": e2e_ok},
            "gpu_launches": args.steps * blk.launches_per_spmv(),
            "clocks": clocks,
            "parity": {"rows_outside_1e-12_gate_all_ranks": int(gate_all), "max_abs_err_rank0": max_err,
                       "checked": "every row of every rank against the closed-form check vector (oracle)"},
        }
        print(json.dumps(out), flush=True)
    for p in pin:
        lib.ehyb_host_free_pinned(p)
    dist.barrier()
    blk.free()
    dec.free()
    dist.barrier()
    if world > 1:
        tdist.destroy_process_group()


def run_reference(args):
    """The reference has no CPU SpMV routine of its own: its CPU path is the CSR
    product over its own arrays (rowIdx/J/V after matrixReorder).  Built here with the oracle's
    restatement of the reader + reorder (the pinned mt-metis binary provides the partition) and
    timed with every host thread.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    # every host thread this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which
    # would time the CPU baseline on one core (must be set before the OpenMP runtime of the oracle loads)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    from oracle import oracle as O
    orc = O.Oracle()
    grid, workload, sample = GRID, WORKLOAD, "%d CSR products of the whole matrix per run" % args.steps
    if args.gpus > 1 and not SLAB and not os.environ.get("EHYB_BENCH_GRID"):
        # our arm's workload at N > 1 is BASELINE.json configs[4] (27-point 512^3, 3.61 G entries: 43 GB as CSR).
        # Bounded sample of it for the host: a 512 x 512 x 8 slab of the same stencil (n 2 097 152, 55.9 M
        # entries) - the CSR rate of a matrix this far above the caches does not depend on how many slabs follow
        grid = (GRID5[0], GRID5[1], 8)
        workload = ("3D 27-point stencil %dx%dx%d fp64 (BASELINE.json configs[4]); bounded sample: a %dx%dx%d slab of it"
                    % (GRID5 + grid))
        sample = "%d CSR products of a %dx%dx%d slab of the 512^3 stencil per run" % ((args.steps,) + grid)
    n, li, lj, lv = O.gen_stencil27_lower(*grid)
    x = orc.x_reference(n)
    m = orc.read_sym(n, li, lj, lv)
    P, W, _ = orc.heuristic_ref(n, True)
    xadj, adj = orc.graph(m)
    part = O.mtmetis_partition(xadj, adj, P, nthreads=1)
    r = orc.reorder(m, P, W, part)
    xr = orc.vector_reorder(x, r["reorderList"])
    sec_w, _ = orc.csr_spmv_timed(r["rowIdx"], r["J"], r["V"], xr, args.warmup, 1)
    sec, _ = orc.csr_spmv_timed(r["rowIdx"], r["J"], r["V"], xr, 0, args.steps)
    gflops = 2.0 * m["nnz"] * args.steps / sec / 1e9
    out = {
        "impl": "reference", "metric": "fp64 SpMV GFLOP/s (2*nnz/t), EHYB format", "value": round(gflops, 3),
        "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(sec / args.steps * 1e3, 4), "higher_is_better": True, "scaling": "strong" if (args.gpus > 1 and not SLAB) or STRONG else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "grid": list(grid), "n": n, "nnz": m["nnz"],
                   "partitions": P, "window": W, "note": "reference partition parameters (82-SM heuristic)"},
        "cpu_baseline": {"value": round(gflops, 3), "unit": "GFLOP/s", "cores": orc.num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(gflops, 3), "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
