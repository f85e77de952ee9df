/*
 * ehyb_kernels.cuh -- device code of the EHYB SpMV engine for sm_100a (B200).
 *
 * Replaces the reference's kernelCachedBlockedELL / _small, vecReorderER and longRowKernel
 * (reference kernel.cu:110-195, :197-284, :69-77, :43-67).  Semantics are the reference's
 * (SURVEY.md A.4): y = A_ell * x(window) + A_rem * x(global); the design is not:
 *
 *   - the x window of a partition is staged into shared memory by TMA bulk copies
 *     (cp.async.bulk.shared::cluster.global + mbarrier complete_tx; SASS: UBLKCP), issued by
 *     one thread, instead of a strided copy loop by all threads;
 *   - values and 16-bit window-local columns stream from HBM with 128-bit loads
 *     (one LDG.128 = two rows of one ELL column; one LDG.128 = two rows of four columns),
 *     bypassing L1 and marked evict-first in L2 so that x stays cached;
 *   - a warp owns a 64-row slice, lane l the rows l and l+32: two independent fp64
 *     accumulator chains per thread, and y is written ONCE, with fully coalesced 256-byte
 *     stores - the in-slice remainder is accumulated by the same lane (no yER round trip, no
 *     scatter-add kernel, no global work counter to reset: SURVEY.md B-1, B-2);
 *   - irregular spill and long rows go through a COO list reduced with warp-shuffle
 *     segmented sums (ehyb_overflow_kernel), one atomic per row segment.
 *
 * No tensor cores: SpMV is not a dense contraction (arithmetic intensity ~0.19 flop/B).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ehyb.h"

namespace ehyb {

constexpr int kSmemHeader = 128; /* mbarrier + slice counter in front of the window */

/*
 * Peer-memory halo exchange (multi-GPU, one process per GPU).  Every GPU
 * owns two halo buffers (products alternate between them) and kMaxPushCtas flag words per rank,
 * all in one allocation that the neighbours map through CUDA IPC.  Product number `epoch`:
 *   push   the last warp of the first pushCtas CTAs of the main kernel stores its share of the
 *          x entries the neighbours need straight into the neighbours' halo buffers over NVLink
 *          and then writes `epoch` into ITS OWN flag word on every neighbour with one
 *          st.release.sys - the only MEMBAR of the protocol (a MEMBAR.SYS drains behind the
 *          ~120 KB the SM has in flight: several microseconds under load, measured);
 *   pull   a warp that needs a halo column (remainder-cache fill, overflow kernel) polls, with
 *          ld.acquire.sys (LDG.STRONG.SYS + CCTL.IVALL, no MEMBAR), until all the flag words of
 *          the neighbours it receives from show `epoch`.
 * flags == NULL: no exchange inside the kernels (single GPU, or the NCCL exchange, where the
 * halo entries live in the overflow list and x_halo = x + n).
 */
constexpr int kMaxPushCtas = 256; /* flag words per rank (>= SMs of the device) */

struct PeerArgs {
    const double *xh;            /* halo values of this product: column c >= n is xh[c - n] */
    const uint32_t *flags;       /* [nranks][kMaxPushCtas] last epoch delivered by CTA j of rank g */
    uint32_t *status;            /* set to 1 when a wait ran into the time limit */
    const int32_t *pushIdx;      /* [pushCount] local x entries the neighbours need */
    double *const *pushDst;      /* [pushCount] their addresses in the neighbours' halo buffers */
    uint32_t *const *peerFlag;   /* [nPeers] address of flags[my rank][0] on every neighbour */
    const int32_t *peerPushCtas; /* [nranks] how many CTAs of rank g push (= flag words to poll) */
    unsigned long long timeoutNs;
    uint32_t epoch;
    uint32_t recvMask;           /* ranks this GPU receives halo values from */
    uint32_t nbrMask;            /* ranks it exchanges flags with (senders and receivers) */
    int nranks, pushCount, pushCtas, nPeers;
};

struct MainArgs {
    const ehyb_part_desc *parts;
    const ehyb_slice_desc *slices;
    const unsigned char *blob;
    const double *x;
    double *y;
    int n;    /* local rows == end of the windowable x range */
    int W;    /* window length in elements */
    int kpp;  /* CTAs per partition */
    int nPartsTotal; /* persistent kernel: partitions of the matrix (its grid is smaller) */
    int dbg;   /* development only (EHYB_DEBUG_SKIP): 1 = skip remainder math, 2 = skip ELL math */
    const int32_t *cacheCols; /* per-partition remainder cache lists (permuted columns) */
    int cacheCap;             /* shared-memory cache capacity in elements (>= longest list) */
    const int32_t *order;     /* CTA slot -> partition (NULL: identity) */
    const int32_t *ctaTab;    /* persistent kernel: work items in CTA order, 8 ints each (see there) */
    const int32_t *ctaStart;  /* persistent kernel: [grid + 1] first work item of every CTA */
    int prologueBarrier;      /* staged kernel, experiments: CTA barrier after window + cache staging */
    int dynamicDeal;          /* staged kernel: slices beyond the first nw are taken on demand (shared-memory counter) */
    int l2hint;               /* staged kernel: L2 eviction hints on the TMA copies (stream evict-first, x evict-last) */
    uint32_t winPiece;        /* staged kernel: bytes per bulk copy of the x window (multiple of 16) */
    unsigned long long *trace; /* development (EHYB_TRACE=1): 8 globaltimer stamps per CTA, else NULL */
    double *dot;              /* ehyb_spmv_dot (DOT builds of the persistent kernel, multi-GPU builds of the staged one):
                                 *dot += sum over the rows of y[r] * x[r], the p.Ap of a CG iteration, x[r] taken
                                 from the x window in shared memory while y is stored */
    PeerArgs peer;
};

struct OverflowArgs {
    const int32_t *row, *col;
    const double *val;
    int64_t count;
    const double *x;
    double *y;
    int perWarp; /* consecutive entries per warp, a multiple of 32 */
    int n;       /* columns >= n are halo columns: peer.xh[c - n] */
    int lateTrigger; /* release the dependent grid at the end of the CTA instead of the top */
    PeerArgs peer;
};

/* ---------------------------------------------------------------- PTX helpers ----- */

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

/* TMA 1-D bulk copy global -> shared, completion counted in bytes on `bar` */
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

/* the same with an L2 eviction-priority hint (createpolicy): evict-first for the matrix stream,
 * which is read once per product, evict-last for x, which every product reads again */
__device__ __forceinline__ void tma_bulk_g2s_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}

__device__ __forceinline__ uint64_t make_evict_last_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

/* streaming 128-bit loads: read-only path, no L1 allocation, evict-first in L2 */
__device__ __forceinline__ uint64_t make_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ double2 ld_stream_f64x2(const double2 *p, uint64_t pol)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ uint4 ld_stream_u32x4(const uint4 *p, uint64_t pol)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ int ld_stream_s32(const int *p, uint64_t pol)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ double ld_stream_f64(const double *p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ int2 ld_stream_s32x2(const int2 *p, uint64_t pol)
{
    int2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.s32 {%0, %1}, [%2], %3;"
                 : "=r"(v.x), "=r"(v.y)
                 : "l"(p), "l"(pol));
    return v;
}

/* x gather of the remainder: a COHERENT load served by L2 (ld.global.cg).  Not the read-only
 * (.nc) path: under programmatic dependent launch this kernel is already resident while the
 * stream's previous kernel still writes x (a solver's vector update), and .nc requires data that
 * is read-only for the whole lifetime of the kernel.  Every gather happens once per CTA. */
__device__ __forceinline__ double ld_gather_f64(const double *p)
{
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

/* ---------------------------------------------------------------- peer exchange --- */

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory"); /* clobber: keeps its place between barriers and memory operations */
    return t;
}

/* Warp-collective (all 32 lanes, converged): waits until the first `firstOnly ? 1 : pushCtas(g)`
 * flag words of every rank g in `mask` show `epoch` or later (flags only grow; compared modulo
 * 2^32).  Bounded (timeoutNs != 0): a peer that does not show up within the limit sets *status
 * - sticky, read by the host after every synchronisation: ehyb_sync, ehyb_get_y and the timed
 * loops then fail with EHYB_ERR_PEER - and the function returns false: the caller must not push
 * into a buffer the neighbour may still read.  The limit therefore also bounds the launch skew
 * between the ranks; timeoutNs == 0 waits without limit. */
__device__ __noinline__ bool peer_wait(const uint32_t *flags, uint32_t mask, int nranks, const int32_t *peerPushCtas, bool firstOnly,
                                       uint32_t epoch, unsigned long long timeoutNs, uint32_t *status)
{
    const int lane = threadIdx.x & 31;
    unsigned long long t0 = 0;
    for (int g = 0; g < nranks; ++g) {
        if (!((mask >> g) & 1u)) continue;
        const uint32_t *f = flags + g * kMaxPushCtas;
        const int cnt = firstOnly ? 1 : __ldg(peerPushCtas + g);
        for (;;) {
            bool ok = true;
            for (int i = lane; i < cnt; i += 32) ok = ok && static_cast<int32_t>(ld_acquire_sys_u32(f + i) - epoch) >= 0;
            if (__all_sync(0xffffffffu, ok)) break;
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (timeoutNs != 0 && __any_sync(0xffffffffu, now - t0 > timeoutNs)) {
                if (lane == 0) atomicExch_system(status, 1u);
                return false;
            }
            __nanosleep(64);
        }
    }
    return true;
}

/* halo value of column c >= n: written by a peer over NVLink, L2 is the point of coherence */
__device__ __forceinline__ double ld_halo_f64(const PeerArgs &pa, int n, int c) { return __ldcg(pa.xh + (c - n)); }

/* Fills the shared-memory remainder cache of a partition from its column list (ascending, so
 * halo columns are its tail).  Own columns are gathered right away; if the exchange runs inside
 * this kernel (pa.flags) and the list has halo columns, warp 0 waits once for the neighbours'
 * push of this product, a CTA barrier orders everybody behind it, and the tail is read from the
 * halo buffer.  Called by all threads of the CTA (direct kernel, unaligned-x path). */
__device__ __forceinline__ void fill_remainder_cache(double *cache, const int32_t *cols, int count, const double *x, int n,
                                                     const PeerArgs &pa, int tid, int nthreads)
{
    bool skipped = false;
    for (int i = tid; i < count; i += nthreads) {
        const int c = __ldg(cols + i);
        if (c < n) cache[i] = ld_gather_f64(x + c);
        else if (pa.flags == nullptr) cache[i] = ld_halo_f64(pa, n, c); /* halo already in place (x tail) */
        else skipped = true;
    }
    if (pa.flags != nullptr && __syncthreads_or(skipped)) {
        if (tid < 32) peer_wait(pa.flags, pa.recvMask, pa.nranks, pa.peerPushCtas, false, pa.epoch, pa.timeoutNs, pa.status);
        __syncthreads();
        for (int i = tid; i < count; i += nthreads) {
            const int c = __ldg(cols + i);
            if (c >= n) cache[i] = ld_halo_f64(pa, n, c);
        }
    }
}

/* The push half, executed by ONE warp (the last one: under the static deal of slices it is a
 * warp with the fewest slices) of every CTA blockIdx.x < pushCtas.  Two steps:
 *   prepare (before griddepcontrol.wait, i.e. hidden when the grid was launched early): check
 *           that the neighbours are done with the halo buffer of this parity, and fetch the
 *           first 128 (index, address) pairs of this CTA's share into registers;
 *   send    (x is final): gather, store to the neighbours, one st.release.sys per neighbour.
 * No thread of a pushing warp waits for this product's flags before it has signalled: two GPUs
 * doing that would deadlock. */
struct PushRegs {
    int idx[4];
    unsigned long long dst[4];
    int i0, i1;
    bool ok; /* the neighbours are done with the halo buffer of this parity */
};

__device__ __forceinline__ void peer_push_prepare(const PeerArgs &pa, int lane, PushRegs &pr)
{
    /* the neighbours' halo buffer of this parity was last read by their product epoch-2, which
     * is complete once any of their CTAs has signalled epoch-1 (it passed its dependency wait) */
    pr.ok = peer_wait(pa.flags, pa.nbrMask, pa.nranks, pa.peerPushCtas, true, pa.epoch - 1u, pa.timeoutNs, pa.status);
    const int chunk = (pa.pushCount + pa.pushCtas - 1) / pa.pushCtas;
    pr.i0 = static_cast<int>(blockIdx.x) * chunk;
    pr.i1 = min(pr.i0 + chunk, pa.pushCount);
    const unsigned long long *dstTab = reinterpret_cast<const unsigned long long *>(pa.pushDst);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = pr.i0 + u * 32 + lane;
        pr.idx[u] = i < pr.i1 ? __ldg(pa.pushIdx + i) : 0;
        pr.dst[u] = i < pr.i1 ? __ldg(dstTab + i) : 0ull;
    }
}

__device__ __forceinline__ void peer_push_send(const PeerArgs &pa, const double *x, int lane, const PushRegs &pr)
{
    /* a neighbour that never released the buffer (time limit, *status set): neither data nor flag -
     * it would overwrite values still being read; the neighbours' waits then fail loudly as well */
    if (!pr.ok) return;
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld_gather_f64(x + pr.idx[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (pr.dst[u]) *reinterpret_cast<double *>(pr.dst[u]) = v[u];
    /* shares beyond 128 entries: same pattern, loads not hoisted above the dependency wait */
    const unsigned long long *dstTab = reinterpret_cast<const unsigned long long *>(pa.pushDst);
    for (int base = pr.i0 + 128; base < pr.i1; base += 128) {
        int idx[4];
        unsigned long long dst[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + u * 32 + lane;
            idx[u] = i < pr.i1 ? __ldg(pa.pushIdx + i) : 0;
            dst[u] = i < pr.i1 ? __ldg(dstTab + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_gather_f64(x + idx[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (dst[u]) *reinterpret_cast<double *>(dst[u]) = v[u];
    }
    __syncwarp(); /* every lane's stores are ordered before lane j's release below */
    if (lane < pa.nPeers) st_release_sys_u32(pa.peerFlag[lane] + blockIdx.x, pa.epoch);
    for (int j = 32 + lane; j < pa.nPeers; j += 32) st_release_sys_u32(pa.peerFlag[j] + blockIdx.x, pa.epoch);
}

/* The same push without registers held across the dependency wait (persistent kernel: its register
 * budget is the consumer loop's): `ok` = result of the release check, made before the wait. */
__device__ __forceinline__ bool peer_push_check(const PeerArgs &pa)
{
    return peer_wait(pa.flags, pa.nbrMask, pa.nranks, pa.peerPushCtas, true, pa.epoch - 1u, pa.timeoutNs, pa.status);
}

__device__ __noinline__ void peer_push_all(const PeerArgs &pa, const double *x, int lane, bool ok)
{
    if (!ok) return; /* see peer_push_send */
    const int chunk = (pa.pushCount + pa.pushCtas - 1) / pa.pushCtas;
    const int i0 = static_cast<int>(blockIdx.x) * chunk, i1 = min(i0 + chunk, pa.pushCount);
    const unsigned long long *dstTab = reinterpret_cast<const unsigned long long *>(pa.pushDst);
    for (int base = i0; base < i1; base += 128) {
        int idx[4];
        unsigned long long dst[4];
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + u * 32 + lane;
            idx[u] = i < i1 ? __ldg(pa.pushIdx + i) : 0;
            dst[u] = i < i1 ? __ldg(dstTab + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_gather_f64(x + idx[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (dst[u]) *reinterpret_cast<double *>(dst[u]) = v[u];
    }
    __syncwarp(); /* every lane's stores are ordered before lane j's release below */
    for (int j = lane; j < pa.nPeers; j += 32) st_release_sys_u32(pa.peerFlag[j] + blockIdx.x, pa.epoch);
}

/* one ELL group: 4 columns x 2 rows per lane */
struct Group {
    uint4 c;
    double2 v0, v1, v2, v3;
};

__device__ __forceinline__ void load_group(Group &g, const uint4 *ec, const double2 *ev, int idx, uint64_t pol)
{
    g.c = ld_stream_u32x4(ec + idx * 32, pol);
    g.v0 = ld_stream_f64x2(ev + (4 * idx + 0) * 32, pol);
    g.v1 = ld_stream_f64x2(ev + (4 * idx + 1) * 32, pol);
    g.v2 = ld_stream_f64x2(ev + (4 * idx + 2) * 32, pol);
    g.v3 = ld_stream_f64x2(ev + (4 * idx + 3) * 32, pol);
}

__device__ __forceinline__ void fma_group(const Group &g, const double *xs, double &acc0, double &acc1)
{
    acc0 = fma(g.v0.x, xs[g.c.x & 0xffffu], acc0);
    acc1 = fma(g.v0.y, xs[g.c.z & 0xffffu], acc1);
    acc0 = fma(g.v1.x, xs[g.c.x >> 16], acc0);
    acc1 = fma(g.v1.y, xs[g.c.z >> 16], acc1);
    acc0 = fma(g.v2.x, xs[g.c.y & 0xffffu], acc0);
    acc1 = fma(g.v2.y, xs[g.c.w & 0xffffu], acc1);
    acc0 = fma(g.v3.x, xs[g.c.y >> 16], acc0);
    acc1 = fma(g.v3.y, xs[g.c.w >> 16], acc1);
}

/* ---------------------------------------------------------------- main kernel ----- */

/*
 * grid  = nParts * kpp CTAs; CTA b serves partition b / kpp and the slices t of that partition
 *         with t % kpp == b % kpp (static interleave - nothing global to reset between launches)
 * block = any multiple of 32 up to 1024
 * smem  = kSmemHeader + align16(W + 2) * 8 + cacheCap * 8 bytes (dynamic)
 */
template <int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) ehyb_main_kernel(const MainArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    int *sliceCounter = reinterpret_cast<int *>(smem + 8);
    double *win = reinterpret_cast<double *>(smem + kSmemHeader);

    const int kpp = a.kpp;
    const int slot = blockIdx.x / kpp;
    const int sub = blockIdx.x - slot * kpp;
    const int p = a.order ? __ldg(a.order + slot) : slot;
    const int4 part = __ldg(reinterpret_cast<const int4 *>(a.parts) + 2 * p);
    const int4 part2 = __ldg(reinterpret_cast<const int4 *>(a.parts) + 2 * p + 1); /* cacheStart, cacheCount */
    const int ps = part.x, pe = part.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    /* multi-GPU: the last warp of the first pushCtas CTAs sends this CTA's share of the halo */
    if (a.peer.flags != nullptr && static_cast<int>(blockIdx.x) < a.peer.pushCtas && warp == nwarps - 1) {
        PushRegs pr;
        peer_push_prepare(a.peer, lane, pr);
        peer_push_send(a.peer, a.x, lane, pr);
    }
    if (pe <= ps) return; /* empty partition (uniform over the CTA) */
    double *cache = win + ((a.W + 2 + 15) & ~15); /* remainder cache behind the window */

    /* ---- stage the x window: x[g0, winEnd) -> win[0, len), g0 = ps rounded down to even so
     *      that the global source is 16-byte aligned; window element c lives at xs[c] ---- */
    const int g0 = ps & ~1;
    const int winEnd = min(ps + a.W, a.n);
    const int len = winEnd - g0;
    const double *xs = win + (ps - g0);
    const bool tma_ok = (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
    const uint32_t barAddr = smem_u32(bar);
    if (tid == 0) {
        *sliceCounter = nwarps;
        mbar_init(barAddr, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tma_ok) {
        if (tid == 0) {
            const uint32_t bulkBytes = static_cast<uint32_t>(len & ~1) * 8u;
            mbar_expect_tx(barAddr, bulkBytes);
            const char *src = reinterpret_cast<const char *>(a.x + g0);
            uint32_t dst = smem_u32(win);
            for (uint32_t off = 0; off < bulkBytes; off += 32768u) {
                const uint32_t chunk = min(32768u, bulkBytes - off);
                tma_bulk_g2s(dst + off, src + off, chunk, barAddr);
            }
        } else if (tid == blockDim.x - 1 && (len & 1)) {
            win[len - 1] = a.x[g0 + len - 1]; /* odd tail element */
        }
    } else {
        for (int i = tid; i < len; i += blockDim.x) win[i] = a.x[g0 + i];
    }
    /* remainder cache: x at the partition's most referenced columns outside the window (halo
     * columns included when the exchange runs inside this kernel) */
    fill_remainder_cache(cache, a.cacheCols + part2.x, part2.y, a.x, a.n, a.peer, tid, blockDim.x);

    /* ---- slices of this CTA: local index t = sub + kpp*q, q handed out dynamically ---- */
    const int nsl = part.w - part.z;
    const int nq = (nsl - sub + kpp - 1) / kpp;
    const uint64_t pol = make_evict_first_policy();
    int q = warp;

    /* first descriptor before waiting for the window (hides its latency) */
    const uint2 *descs = reinterpret_cast<const uint2 *>(a.slices) + part.z + sub; /* {off256, w | wr<<16} */
    uint2 d = make_uint2(0u, 0u);
    if (q < nq) d = __ldg(descs + kpp * q);

    if (tma_ok) {
        while (!mbar_try_wait(barAddr, 0)) { }
    }
    __syncthreads(); /* tail element / fallback copy visible */

    while (q < nq) {
        const int t = sub + kpp * q;
        const unsigned char *base = a.blob + static_cast<size_t>(d.x) * 256u;
        const int w = d.y & 0xffffu, wr = d.y >> 16;
        const double2 *ev = reinterpret_cast<const double2 *>(base) + lane;
        const uint4 *ec = reinterpret_cast<const uint4 *>(base + static_cast<size_t>(w) * 512u) + lane;
        double acc0 = 0.0, acc1 = 0.0;

        /* ELL part, software pipelined by hand: the loads of group g+1 (one 128-bit column
         * load + four 128-bit value loads = 2.5 KB per warp) are issued before group g is
         * consumed, so a warp always has a full group in flight while it gathers and
         * multiplies.  (Left to itself the compiler interleaves loads and their first uses and
         * a warp ends up waiting on memory three times per group with < 2 KB outstanding.) */
        const int nfull = w >> 2;
        Group ga, gb;
        if (nfull > 0) load_group(ga, ec, ev, 0, pol);
        int g = 0;
        for (; g + 2 <= nfull; g += 2) {
            load_group(gb, ec, ev, g + 1, pol);
            fma_group(ga, xs, acc0, acc1);
            if (g + 2 < nfull) load_group(ga, ec, ev, g + 2, pol);
            fma_group(gb, xs, acc0, acc1);
        }
        if (g < nfull) fma_group(ga, xs, acc0, acc1);
        const int tail = w & 3;
        if (tail) {
            const uint4 c = ld_stream_u32x4(ec + nfull * 32, pol);
            const double2 v0 = ld_stream_f64x2(ev + (4 * nfull + 0) * 32, pol);
            double2 v1 = make_double2(0.0, 0.0), v2 = make_double2(0.0, 0.0);
            if (tail > 1) v1 = ld_stream_f64x2(ev + (4 * nfull + 1) * 32, pol);
            if (tail > 2) v2 = ld_stream_f64x2(ev + (4 * nfull + 2) * 32, pol);
            acc0 = fma(v0.x, xs[c.x & 0xffffu], acc0);
            acc1 = fma(v0.y, xs[c.z & 0xffffu], acc1);
            if (tail > 1) {
                acc0 = fma(v1.x, xs[c.x >> 16], acc0);
                acc1 = fma(v1.y, xs[c.z >> 16], acc1);
            }
            if (tail > 2) {
                acc0 = fma(v2.x, xs[c.y & 0xffffu], acc0);
                acc1 = fma(v2.y, xs[c.w & 0xffffu], acc1);
            }
        }

        /* in-slice remainder: same geometry as the ELL part, 16-bit indices into the
         * remainder cache, own accumulators (y = dot_ell + dot_rem, kernel.cu:162 + :76) */
        if (wr) {
            const size_t remOff = static_cast<size_t>(w) * 512u + static_cast<size_t>((w + 3) >> 2) * 512u;
            const double2 *rv = reinterpret_cast<const double2 *>(base + remOff) + lane;
            const uint4 *rc = reinterpret_cast<const uint4 *>(base + remOff + static_cast<size_t>(wr) * 512u) + lane;
            double r0 = 0.0, r1 = 0.0;
            const int rfull = wr >> 2;
            for (int q2 = 0; q2 < rfull; ++q2) {
                Group gr;
                load_group(gr, rc, rv, q2, pol);
                fma_group(gr, cache, r0, r1);
            }
            const int rtail = wr & 3;
            if (rtail) {
                const uint4 c = ld_stream_u32x4(rc + rfull * 32, pol);
                const uint32_t c0[3] = {c.x & 0xffffu, c.x >> 16, c.y & 0xffffu};
                const uint32_t c1[3] = {c.z & 0xffffu, c.z >> 16, c.w & 0xffffu};
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (i < rtail) {
                        const double2 v = ld_stream_f64x2(rv + (4 * rfull + i) * 32, pol);
                        r0 = fma(v.x, cache[c0[i]], r0);
                        r1 = fma(v.y, cache[c1[i]], r1);
                    }
                }
            }
            acc0 += r0;
            acc1 += r1;
        }

        /* next slice + its descriptor while the stores drain */
        int qn = 0;
        if (lane == 0) qn = atomicAdd(sliceCounter, 1);
        qn = __shfl_sync(0xffffffffu, qn, 0);
        uint2 dn = make_uint2(0u, 0u);
        if (qn < nq) dn = __ldg(descs + kpp * qn);

        const int r = ps + t * EHYB_SLICE_ROWS + lane;
        if (r < pe) a.y[r] = acc0;
        if (r + 32 < pe) a.y[r + 32] = acc1;
        q = qn;
        d = dn;
    }
}

/* Warp-level variant for the staged kernel: warp `fw` of `nFill` filling warps takes every
 * nFill-th group of 32 list entries; a warp that meets halo columns waits for the neighbours'
 * push itself (one poll per warp).  No CTA barrier: the caller arrives on an mbarrier. */
__device__ __forceinline__ void fill_remainder_cache_warp(double *cache, const int32_t *cols, int count, const double *x, int n,
                                                          const PeerArgs &pa, int fw, int nFill, int lane)
{
    bool skipped = false;
    for (int i = fw * 32 + lane; i < count; i += nFill * 32) {
        const int c = __ldg(cols + i);
        if (c < n) cache[i] = ld_gather_f64(x + c);
        else if (pa.flags == nullptr) cache[i] = ld_halo_f64(pa, n, c);
        else skipped = true;
    }
    if (pa.flags != nullptr && __any_sync(0xffffffffu, skipped)) {
        peer_wait(pa.flags, pa.recvMask, pa.nranks, pa.peerPushCtas, false, pa.epoch, pa.timeoutNs, pa.status);
        for (int i = fw * 32 + lane; i < count; i += nFill * 32) {
            const int c = __ldg(cols + i);
            if (c >= n) cache[i] = ld_halo_f64(pa, n, c);
        }
    }
}

/* ---------------------------------------------------------------- staged kernel --- */

/*
 * Same product, but the matrix stream itself goes through shared memory: every warp owns
 * kSlotsPerWarp staging slots and keeps them filled with TMA bulk copies of the next chunks
 * of its slices (a chunk = up to 16 ELL columns, or up to 8 remainder columns, of one 64-row
 * slice: 10 KB).  The bytes in flight per SM are then set by the staging capacity
 * (warps x slots x 10 KB, ~100 KB next to a 114 KB window) instead of by the registers a
 * thread can devote to outstanding loads - ncu showed the direct kernel holding only ~24 KB
 * per SM in flight and HBM 55 % busy.  Consumers read values (LDS.128), columns (LDS.128 /
 * LDS.64) and gather x from the window (LDS.64).
 *
 * grid  = nParts * kpp, block = NW*32 threads,
 * smem  = kStageHeader + align128((W+2)*8) + NW * kSlotsPerWarp * kSlotBytes
 * Slices are dealt statically: warp i of CTA `sub` takes local slices sub + kpp*(i + NW*j).
 */
constexpr int kStageHeader = 512;  /* window + cache mbarriers, 2*24 slot mbarriers, slice counter (last 8 bytes) */
constexpr int kSlotsPerWarp = 2;
constexpr int kMaxStageWarps = 24;
/* a slot holds one chunk: kc columns of one slice, ELL or remainder alike (kc*512 B of values
 * followed at slot_val_bytes() by kc/4*512 B of 16-bit indices) */
__host__ __device__ constexpr int slot_val_bytes(int kc) { return kc * 512; }
__host__ __device__ constexpr int slot_bytes(int kc) { return kc * 640; }

__device__ __forceinline__ double2 lds_f64x2(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u32x4(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int2 lds_s32x2(uint32_t addr)
{
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

/* Fused dot product of a CG iteration (MainArgs.dot): rows r and r + 32 of the slice just stored,
 * y[r] * x[r] with x[r] taken from the partition's window in shared memory (rows of a partition
 * start at the window's first element; a row beyond the window reads x from global memory). */
__device__ __forceinline__ double dot_rows(const double *x, uint32_t xsAddr, int ps, int winEnd, int pe, int r, double y0, double y1, double dacc)
{
    if (r < pe) dacc = fma(y0, r < winEnd ? lds_f64(xsAddr + static_cast<uint32_t>(r - ps) * 8u) : ld_gather_f64(x + r), dacc);
    if (r + 32 < pe) dacc = fma(y1, r + 32 < winEnd ? lds_f64(xsAddr + static_cast<uint32_t>(r + 32 - ps) * 8u) : ld_gather_f64(x + r + 32), dacc);
    return dacc;
}
/* end of the warp: one atomic per warp (the sum order over warps is not fixed: like every reduction
 * of the solver, inside its tolerance, not bit-reproducible) */
__device__ __forceinline__ void dot_flush(double *dot, double dacc, int lane)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, off);
    if (lane == 0) atomicAdd(dot, dacc);
}

/* shared-memory address of x[idx] for the two 16-bit indices packed in c: base + 8 * index.  Written
 * as extract + multiply-add (2 instructions per gather; left to itself the compiler shifts, masks
 * and adds: 3) - the index arithmetic is a third of the consumer loop's instructions. */
__device__ __forceinline__ uint32_t xaddr_lo(uint32_t c, uint32_t base)
{
    uint32_t r;
    asm("{\n.reg .u32 t;\nand.b32 t, %1, 0xffff;\nmad.lo.u32 %0, t, 8, %2;\n}" : "=r"(r) : "r"(c), "r"(base));
    return r;
}
__device__ __forceinline__ uint32_t xaddr_hi(uint32_t c, uint32_t base)
{
    uint32_t r;
    asm("{\n.reg .u32 t;\nshr.u32 t, %1, 16;\nmad.lo.u32 %0, t, 8, %2;\n}" : "=r"(r) : "r"(c), "r"(base));
    return r;
}

/* Walks the chunks of the slices of one warp; every member is warp-uniform.  Chunk sizes are
 * compile-time powers of two so that the per-chunk bookkeeping is a handful of shifts (ncu on
 * the first version showed the consumers spending most of their issue slots on this
 * bookkeeping, integer divisions included, rather than on the matrix entries). */
template <int KCE>
struct ChunkWalker {
    const uint2 *descs;        /* descriptors of this CTA's slices: descs[q * kpp], q = 0..nq-1 */
    const unsigned char *blob;
    const unsigned char *base; /* current slice */
    int *counter;              /* shared-memory counter dealing the slices q >= nw (NULL: static deal) */
    int kpp, sub, nq, nw;
    int q, qnext;              /* current and next slice of this warp (index among the CTA's slices) */
    int t;                     /* current slice (local index in the partition) */
    int w, wr, nE, nc, ci;     /* current slice: widths, ELL chunks, all chunks, next chunk */
    uint2 dnext;               /* prefetched descriptor of the next slice */
    bool live;

    __device__ __forceinline__ void load_slice(uint2 d)
    {
        base = blob + static_cast<size_t>(d.x) * 256u;
        w = static_cast<int>(d.y & 0xffffu);
        wr = static_cast<int>(d.y >> 16);
        nE = (w + KCE - 1) / KCE;
        nc = max(1, nE + (wr + KCE - 1) / KCE);
        ci = 0;
    }

    /* the slice after `cur` for this warp: the next one nobody has taken (dynamic: the rows of a
     * partition are sorted by length, so slices come widest first and the deal is longest-job-
     * first), or cur + nw (static) */
    __device__ __forceinline__ int take_next(int cur, int lane)
    {
        if (counter == nullptr) return cur + nw;
        int v = 0;
        if (lane == 0) v = atomicAdd(counter, 1);
        return __shfl_sync(0xffffffffu, v, 0);
    }

    __device__ __forceinline__ void start(const uint2 *partDescs, const unsigned char *blob_, int sub_, int kpp_, int warp, int nw_, int nsl,
                                          int *counter_, int lane)
    {
        blob = blob_;
        counter = counter_;
        kpp = kpp_; sub = sub_; nw = nw_;
        nq = nsl > sub ? (nsl - sub + kpp - 1) / kpp : 0;
        descs = partDescs + sub;
        q = warp; /* the first nw slices are dealt statically */
        live = q < nq;
        t = sub + kpp * q;
        qnext = nq;
        dnext = make_uint2(0u, 0u);
        if (live) {
            load_slice(__ldg(descs + kpp * q));
            qnext = take_next(q, lane);
            if (qnext < nq) dnext = __ldg(descs + kpp * qnext);
        }
    }

    __device__ __forceinline__ void advance(int lane)
    {
        if (qnext < nq) {
            q = qnext;
            t = sub + kpp * q;
            load_slice(dnext);
            qnext = take_next(q, lane);
            if (qnext < nq) dnext = __ldg(descs + kpp * qnext);
        } else {
            live = false;
        }
    }
};

struct ChunkMeta {
    int kc;    /* columns in the chunk (0: empty slice) */
    int flags; /* 1 = remainder chunk, 2 = last chunk of its slice, 4 = valid */
    int t;     /* slice the chunk belongs to */
};

/* Describes the next chunk, advances the walker and (lane 0) starts its TMA copies. */
template <int KCE>
__device__ __forceinline__ ChunkMeta issue_chunk(ChunkWalker<KCE> &wk, uint32_t slotAddr, uint32_t barAddr, int lane, uint64_t streamPolicy)
{
    constexpr uint32_t kValBytes = static_cast<uint32_t>(slot_val_bytes(KCE));
    ChunkMeta m;
    m.kc = 0; m.flags = 0; m.t = wk.t;
    if (!wk.live) return m;
    const unsigned char *src0, *src1;
    uint32_t b0, b1;
    m.flags = 4;
    if (wk.ci < wk.nE) {
        const int k = wk.ci * KCE;
        m.kc = min(KCE, wk.w - k);
        src0 = wk.base + static_cast<uint32_t>(k) * 512u;
        b0 = static_cast<uint32_t>(m.kc) * 512u;
        src1 = wk.base + static_cast<uint32_t>(wk.w) * 512u + static_cast<uint32_t>(k >> 2) * 512u;
        b1 = static_cast<uint32_t>((m.kc + 3) >> 2) * 512u;
    } else {
        const int k = (wk.ci - wk.nE) * KCE;
        const unsigned char *rem = wk.base + static_cast<uint32_t>(wk.w) * 512u + static_cast<uint32_t>((wk.w + 3) >> 2) * 512u;
        m.kc = max(0, min(KCE, wk.wr - k)); /* 0 only for a slice without any entry */
        m.flags |= 1;
        src0 = rem + static_cast<uint32_t>(k) * 512u;
        b0 = static_cast<uint32_t>(m.kc) * 512u;
        src1 = rem + static_cast<uint32_t>(wk.wr) * 512u + static_cast<uint32_t>(k >> 2) * 512u;
        b1 = static_cast<uint32_t>((m.kc + 3) >> 2) * 512u;
    }
    if (lane == 0 && b0) {
        mbar_expect_tx(barAddr, b0 + b1);
        if (streamPolicy) {
            tma_bulk_g2s_hint(slotAddr, src0, b0, barAddr, streamPolicy);
            tma_bulk_g2s_hint(slotAddr + kValBytes, src1, b1, barAddr, streamPolicy);
        } else {
            tma_bulk_g2s(slotAddr, src0, b0, barAddr);
            tma_bulk_g2s(slotAddr + kValBytes, src1, b1, barAddr);
        }
    }
    if (++wk.ci == wk.nc) { /* last chunk of the slice: move on */
        m.flags |= 2;
        wk.advance(lane);
    }
    return m;
}

/* PEER: the multi-GPU build of the kernel (halo push, halo columns in the cache fill, per-CTA
 * trace).  Single-GPU sessions launch the PEER = false build, which is ~1 000 instructions
 * shorter: on kernels whose CTAs live a few microseconds the code size is measurable. */
template <int kMaxThreads, int KCE, bool PEER>
__global__ void __launch_bounds__(kMaxThreads, 1) ehyb_staged_kernel(const MainArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr uint32_t kSlotBytes = static_cast<uint32_t>(slot_bytes(KCE));
    constexpr uint32_t kSlotValBytes = static_cast<uint32_t>(slot_val_bytes(KCE));
    const int kpp = a.kpp;
    const int slot_ = blockIdx.x / kpp;
    const int sub = blockIdx.x - slot_ * kpp;
    /* CTAs are dispatched in blockIdx order; `order` lets the session run the partitions that
     * depend on halo values last, when the neighbours have long delivered them */
    const int p = a.order ? __ldg(a.order + slot_) : slot_;
    const int4 part = __ldg(reinterpret_cast<const int4 *>(a.parts) + 2 * p);
    const int4 part2 = __ldg(reinterpret_cast<const int4 *>(a.parts) + 2 * p + 1); /* cacheStart, cacheCount */
    const int ps = part.x, pe = part.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = blockDim.x >> 5;
    const bool pusher = PEER && a.peer.flags != nullptr && static_cast<int>(blockIdx.x) < a.peer.pushCtas && warp == nw - 1;
    unsigned long long *tr = PEER && a.trace ? a.trace + static_cast<size_t>(blockIdx.x) * 8 : nullptr;
    if (tr && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tr[0] = global_timer_ns(); /* CTA start */
        tr[5] = smid;
        tr[6] = static_cast<unsigned long long>(p);
        tr[4] = 0;
        tr[2] = 0;
    }
    PushRegs pr;
    if (pe <= ps) { /* empty partition: only its share of the halo push */
        if (pusher) {
            peer_push_prepare(a.peer, lane, pr);
            asm volatile("griddepcontrol.wait;" ::: "memory");
            peer_push_send(a.peer, a.x, lane, pr);
        }
        return;
    }
    const int g0 = ps & ~1;
    const int winEnd = min(ps + a.W, a.n);
    const int len = winEnd - g0;
    double *win = reinterpret_cast<double *>(smem + kStageHeader);
    const uint32_t winBytes = (static_cast<uint32_t>(a.W + 2) * 8u + 127u) & ~127u;
    const uint32_t xsAddr = smem_u32(win) + static_cast<uint32_t>(ps - g0) * 8u;
    const uint32_t winBar = smem_u32(smem);       /* x window staged (TMA bytes + the odd tail element) */
    const uint32_t cacheBar = smem_u32(smem + 8); /* remainder cache staged (one arrival per filling warp) */
    const uint32_t slotBar0 = smem_u32(smem + 16) + static_cast<uint32_t>(warp * kSlotsPerWarp) * 8u;
    const uint32_t cacheBytes = (static_cast<uint32_t>(a.cacheCap) * 8u + 127u) & ~127u;
    double *cache = reinterpret_cast<double *>(smem + kStageHeader + winBytes);
    const uint32_t cacheAddr = smem_u32(cache);
    const uint32_t slot0 = cacheAddr + cacheBytes + static_cast<uint32_t>(warp * kSlotsPerWarp) * kSlotBytes;
    const bool tma_ok = (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
    /* in a CTA that pushes halo values the last warp does only that; the others fill the cache */
    const bool pushCta = PEER && a.peer.flags != nullptr && static_cast<int>(blockIdx.x) < a.peer.pushCtas;
    const int nFill = pushCta && nw > 1 ? nw - 1 : nw;

    /* Programmatic dependent launch: let the next grid in the stream start as soon as SMs free
     * up.  Its matrix stream (constant data) then overlaps this grid's tail; everything that
     * touches x or y comes after its own griddepcontrol.wait below. */
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int *sliceCounter = reinterpret_cast<int *>(smem + kStageHeader - 8);
    if (tid == 0) {
        *sliceCounter = nw; /* slices 0..nw-1 go to the warps statically, the rest on demand */
        mbar_init(winBar, 2);
        mbar_init(cacheBar, static_cast<uint32_t>(nFill));
        for (int i = 0; i < nw * kSlotsPerWarp; ++i) mbar_init(smem_u32(smem + 16) + i * 8u, 1);
        fence_mbar_init();
    }
    __syncthreads();

    /* the matrix stream does not depend on x, y or the previous grid: start it right away */
    ChunkWalker<KCE> wk;
    wk.start(reinterpret_cast<const uint2 *>(a.slices) + part.z, a.blob, sub, kpp, warp, nw, part.w - part.z,
             a.dynamicDeal ? sliceCounter : nullptr, lane);
    ChunkMeta meta[2];
    /* L2 priorities (a.l2hint): the matrix is read once per product -> evict-first; x is read
     * again by every product -> evict-last, so that 600 MB of stream do not push it out */
    const uint64_t streamPolicy = a.l2hint ? make_evict_first_policy() : 0ull;
    meta[0] = issue_chunk(wk, slot0, slotBar0, lane, streamPolicy);
    meta[1] = issue_chunk(wk, slot0 + kSlotBytes, slotBar0 + 8u, lane, streamPolicy);
    uint32_t phases = 0; /* bit s = parity to wait for on slot s */

    /* multi-GPU, pushing warp: everything of the halo push that does not need x */
    if (pusher) peer_push_prepare(a.peer, lane, pr);

    /* x (and later y) belong to the stream's previous work: wait for it here (every thread
     * reads x below: the window tail / fallback copy and the remainder cache) */
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tr && tid == 0) tr[1] = global_timer_ns(); /* previous grid complete */
    /* The prologue has no CTA-wide barrier: a warp starts on its ELL chunks as soon as the
     * window has landed (winBar) and needs the remainder cache (cacheBar) only at its first
     * remainder chunk; nobody waits for the warp that pushes halo values. */
    bool cacheReady = false;
    if (tma_ok) {
        if (tid == 0) {
            const uint32_t bulkBytes = static_cast<uint32_t>(len & ~1) * 8u;
            mbar_expect_tx(winBar, bulkBytes); /* arrival 1 of 2 */
            const char *src = reinterpret_cast<const char *>(a.x + g0);
            const uint32_t dst = smem_u32(win);
            const uint32_t piece = a.winPiece;
            if (a.l2hint) {
                const uint64_t keep = make_evict_last_policy();
                for (uint32_t off = 0; off < bulkBytes; off += piece) tma_bulk_g2s_hint(dst + off, src + off, min(piece, bulkBytes - off), winBar, keep);
            } else {
                for (uint32_t off = 0; off < bulkBytes; off += piece) tma_bulk_g2s(dst + off, src + off, min(piece, bulkBytes - off), winBar);
            }
        } else if (tid == blockDim.x - 1) {
            if (len & 1) win[len - 1] = a.x[g0 + len - 1]; /* odd tail element */
            mbar_arrive(winBar);                           /* arrival 2 of 2 (releases the store) */
        }
        /* multi-GPU: this CTA's share of the x entries the neighbours need goes out over NVLink
         * (last warp only; it joins the product when it is done) */
        if (pusher) {
            peer_push_send(a.peer, a.x, lane, pr);
            if (tr && lane == 0) tr[2] = global_timer_ns(); /* this CTA's share pushed and signalled */
        }
        /* remainder cache: x at the partition's most referenced columns outside the window,
         * gathered once per CTA (ascending list: neighbouring lanes mostly share sectors; halo
         * columns, >= n, are its tail and wait for the neighbours' push) */
        if (!pusher || nFill == nw) {
            if constexpr (PEER) {
                fill_remainder_cache_warp(cache, a.cacheCols + part2.x, part2.y, a.x, a.n, a.peer, warp, nFill, lane);
            } else { /* no exchange in this build: halo columns, if any, are the tail of x */
                for (int i = warp * 32 + lane; i < part2.y; i += nw * 32) cache[i] = ld_gather_f64(a.x + __ldg(a.cacheCols + part2.x + i));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(cacheBar);
        }
        while (!mbar_try_wait(winBar, 0)) { }
        if (a.prologueBarrier) { /* (experiment switch: the classic CTA barrier behind the staging) */
            __syncthreads();
            cacheReady = true;
        }
    } else {
        /* x not 16-byte aligned: plain copies and a CTA barrier */
        for (int i = tid; i < len; i += blockDim.x) win[i] = a.x[g0 + i];
        if (pusher) peer_push_send(a.peer, a.x, lane, pr);
        if constexpr (PEER) {
            fill_remainder_cache(cache, a.cacheCols + part2.x, part2.y, a.x, a.n, a.peer, tid, blockDim.x);
        } else {
            for (int i = tid; i < part2.y; i += blockDim.x) cache[i] = ld_gather_f64(a.x + __ldg(a.cacheCols + part2.x + i));
        }
        __syncthreads();
        cacheReady = true;
    }
    if (tr && tid == 0) tr[3] = global_timer_ns(); /* x window in shared memory */

    double acc0 = 0.0, acc1 = 0.0, r0 = 0.0, r1 = 0.0;
    double dacc = 0.0; /* PEER build with a.dot: this warp's share of y.x */
    int s = 0; /* slot in use: chunks alternate between the two slots of the warp */
#pragma unroll 1
    while (meta[0].flags | meta[1].flags) {
        const ChunkMeta m = s ? meta[1] : meta[0];
        if (!(m.flags & 4)) break;
        const uint32_t slot = slot0 + static_cast<uint32_t>(s) * kSlotBytes;
        const uint32_t bar = slotBar0 + static_cast<uint32_t>(s) * 8u;
        if (m.kc) {
            while (!mbar_try_wait(bar, (phases >> s) & 1u)) { }
            phases ^= 1u << s;
            if ((m.flags & 1) && !cacheReady) { /* first remainder chunk of this warp */
                while (!mbar_try_wait(cacheBar, 0)) { }
                cacheReady = true;
                if (tr && tid == 0) tr[7] = global_timer_ns(); /* remainder cache in shared memory */
            }
            const uint32_t vAddr = slot + static_cast<uint32_t>(lane) * 16u;
            if (!(a.dbg & ((m.flags & 1) ? 1 : 2))) { /* (dbg: timing experiments skip the arithmetic) */
                /* ELL chunk: indices into the x window, accumulators acc0/acc1; remainder
                 * chunk: indices into the remainder cache, accumulators r0/r1 */
                const bool rem = (m.flags & 1) != 0;
                const uint32_t xb = rem ? cacheAddr : xsAddr;
                double s0 = rem ? r0 : acc0, s1 = rem ? r1 : acc1;
                const uint32_t cAddr = slot + kSlotValBytes + static_cast<uint32_t>(lane) * 16u;
                const int nfull = m.kc >> 2;
#pragma unroll 2
                for (int g = 0; g < nfull; ++g) {
                    const uint4 c = lds_u32x4(cAddr + g * 512u);
                    const double2 v0 = lds_f64x2(vAddr + (4 * g + 0) * 512u);
                    const double2 v1 = lds_f64x2(vAddr + (4 * g + 1) * 512u);
                    const double2 v2 = lds_f64x2(vAddr + (4 * g + 2) * 512u);
                    const double2 v3 = lds_f64x2(vAddr + (4 * g + 3) * 512u);
                    const double x00 = lds_f64(xb + (c.x & 0xffffu) * 8u), x01 = lds_f64(xb + (c.z & 0xffffu) * 8u);
                    const double x10 = lds_f64(xb + (c.x >> 16) * 8u), x11 = lds_f64(xb + (c.z >> 16) * 8u);
                    const double x20 = lds_f64(xb + (c.y & 0xffffu) * 8u), x21 = lds_f64(xb + (c.w & 0xffffu) * 8u);
                    const double x30 = lds_f64(xb + (c.y >> 16) * 8u), x31 = lds_f64(xb + (c.w >> 16) * 8u);
                    s0 = fma(v0.x, x00, s0);
                    s1 = fma(v0.y, x01, s1);
                    s0 = fma(v1.x, x10, s0);
                    s1 = fma(v1.y, x11, s1);
                    s0 = fma(v2.x, x20, s0);
                    s1 = fma(v2.y, x21, s1);
                    s0 = fma(v3.x, x30, s0);
                    s1 = fma(v3.y, x31, s1);
                }
                const int tail = m.kc & 3;
                if (tail) {
                    const uint4 c = lds_u32x4(cAddr + nfull * 512u);
                    const uint32_t cols0[3] = {c.x & 0xffffu, c.x >> 16, c.y & 0xffffu};
                    const uint32_t cols1[3] = {c.z & 0xffffu, c.z >> 16, c.w & 0xffffu};
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        if (i < tail) {
                            const double2 v = lds_f64x2(vAddr + (4 * nfull + i) * 512u);
                            s0 = fma(v.x, lds_f64(xb + cols0[i] * 8u), s0);
                            s1 = fma(v.y, lds_f64(xb + cols1[i] * 8u), s1);
                        }
                    }
                }
                if (rem) { r0 = s0; r1 = s1; } else { acc0 = s0; acc1 = s1; }
            }
        }
        if (m.flags & 2) {
            const int r = ps + m.t * EHYB_SLICE_ROWS + lane;
            if (r < pe) a.y[r] = acc0 + r0; /* y = dot_ell + dot_rem, as kernel.cu:162 + :76 */
            if (r + 32 < pe) a.y[r + 32] = acc1 + r1;
            if constexpr (PEER) {
                if (a.dot != nullptr) dacc = dot_rows(a.x, xsAddr, ps, winEnd, pe, r, acc0 + r0, acc1 + r1, dacc);
            }
            acc0 = acc1 = r0 = r1 = 0.0;
        }
        __syncwarp(); /* every lane is done with the slot before it is refilled */
        const ChunkMeta mn = issue_chunk(wk, slot, bar, lane, streamPolicy);
        if (s) meta[1] = mn; else meta[0] = mn;
        s ^= 1;
    }
    if constexpr (PEER) {
        if (a.dot != nullptr) dot_flush(a.dot, dacc, lane);
    }
    if (tr && lane == 0) atomicMax(tr + 4, global_timer_ns()); /* last warp of the CTA done */
}

/* ---------------------------------------------------------------- persistent kernel --- */

/*
 * The staged kernel stops the matrix stream of an SM for ~6 us at every partition boundary (CTA
 * exit + start, descriptor chain, window + cache staging: per-CTA timeline in
 * profiles/r1_notes.md), while HBM delivers ~7.5 TB/s in the steady phase.  This variant keeps
 * ONE CTA per SM alive over ALL of its partitions and double-buffers the explicit cache:
 *
 *   - shared memory holds two {x window, remainder cache} buffers; while the warps consume
 *     partition j from buffer j&1, partition j+1 is staged into the other one: the window by a
 *     TMA bulk copy (warp 0), the cache by every warp's share of the gathers, both right after a
 *     warp's first slice of partition j ("duty"), behind the `empty` mbarrier that tells that
 *     every warp has left partition j-1;
 *   - the slices of a CTA's partitions form ONE sequence, dealt to the warps from one
 *     shared-memory counter (within a partition the rows are sorted by length, so this is
 *     longest-job-first); a warp's chunk stream does not know partition boundaries: when the
 *     next slice it is dealt lies in a later partition, a "switch" marker per boundary goes down
 *     its two-slot pipeline where the consumer has to change buffers;
 *   - the work of a CTA is a list of ITEMS in GLOBAL memory (a.ctaTab, 32 bytes each, CTA c owns
 *     items ctaStart[c] .. ctaStart[c+1]), an item = a run of consecutive slices of one partition.
 *     The session cuts the sequence of all slices into one run per CTA of equal BYTES, so the
 *     CTAs finish together whatever the number and the sizes of the partitions (a whole number of
 *     partitions per CTA left 48 of 148 SMs idle for the last 1/28 of a 4 096-partition product);
 *     a CTA's first and last items are parts of partitions it shares with its neighbours.  Items
 *     are read through L1/L2 one slice ahead of their use: any number per CTA (27-point 512^3 on
 *     one GPU: 32 768 partitions, 222 per CTA);
 *   - no CTA-wide barrier after the start-up; warps are at most one partition apart.
 *
 * PEER: the multi-GPU build.  The last warp of the first pushCtas CTAs (all of them are
 * resident: the grid is one CTA per SM) sends its share of the x entries the neighbours need
 * before it joins the product; halo columns are the tail of the partitions' ascending cache
 * lists: a warp whose share of a list reaches into the halo waits (once per product) for the
 * neighbours' flags and reads the halo buffer; the session orders every CTA's partitions so that
 * those with halo columns come last, when the neighbours' push has long arrived.
 *
 * grid = min(#SMs, nSlices), block = NW*32, one CTA per SM.
 * smem = 1 KB header (mbarriers, sequence counter) + 2 * (align128((W+2)*8) +
 * align128(cacheCap*8)) + NW * NS * slot.  Requires ctasPerPart == 1.
 */
/* mbarrier wait with a time limit: a protocol error between the warps must abort the kernel
 * (trap: the launch fails with an error), never leave the GPU spinning */
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = global_timer_ns();
    unsigned polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++polls & 1023u) == 0 && global_timer_ns() - t0 > 20000000000ull) __trap();
    }
}

/* asynchronous 8-byte copy global -> shared (LDGSTS) and "arrive on the mbarrier when all my
 * earlier cp.async have landed" (counts as one of the barrier's expected arrivals) */
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}

constexpr int kPersistHeader = 1024;

struct PMeta {
    int kc;    /* columns in the chunk */
    int flags; /* 1 remainder chunk, 2 last chunk of its slice, 4 valid, 8 switch to the next partition (no data) */
    int t;     /* slice (local index in its partition) */
};

/* issue side of a warp: walks the slices it is dealt out of the CTA's sequence */
template <int KCE>
struct PWalker {
    const int4 *tab;           /* this CTA's work items: item j at tab[2 * j] */
    int *counter;              /* shared: next undealt slice of the CTA's sequence */
    const uint2 *slices;       /* global slice descriptors */
    const unsigned char *blob;
    const unsigned char *base; /* current slice */
    int nj;
    int j;                     /* partition of the current slice */
    int t;                     /* current slice, local index in partition j */
    int qn, jn;                /* next slice of this warp (sequence number) and its partition */
    int baseN, nslN, sliceStartN; /* partition jn: first sequence number, slices, first descriptor */
    int w, wr, nE, nc, ci;
    uint2 dnext;
    int switchesOwed;          /* partition boundaries crossed but not yet sent down the pipeline */
    bool live;                 /* a current slice exists */

    __device__ __forceinline__ void load_slice(uint2 d)
    {
        base = blob + static_cast<size_t>(d.x) * 256u;
        w = static_cast<int>(d.y & 0xffffu);
        wr = static_cast<int>(d.y >> 16);
        nE = (w + KCE - 1) / KCE;
        nc = max(1, nE + (wr + KCE - 1) / KCE);
        ci = 0;
    }

    /* takes the next sequence number and finds its partition and descriptor (one slice ahead of
     * their use: the table and descriptor loads overlap the chunks of the current slice) */
    __device__ __forceinline__ void take(int lane)
    {
        int v = 0;
        if (lane == 0) v = atomicAdd(counter, 1);
        qn = __shfl_sync(0xffffffffu, v, 0);
        while (jn < nj && qn >= baseN + nslN) {
            baseN += nslN;
            jn += 1;
            if (jn < nj) {
                const int4 e = __ldg(tab + 2 * jn);
                sliceStartN = e.z;
                nslN = e.w - e.z;
            }
        }
        if (jn < nj) dnext = __ldg(slices + sliceStartN + (qn - baseN));
    }

    /* positions the walker on the next slice this warp gets, crossing partitions as needed; when
     * the sequence is exhausted the remaining boundaries are still owed: every warp passes every
     * partition of the CTA (it owes each of them its share of the staging and an `empty` arrival) */
    __device__ __forceinline__ void next_slice(int lane)
    {
        if (jn >= nj) {
            switchesOwed += nj - 1 - j;
            j = nj - 1;
            live = false;
            return;
        }
        switchesOwed += jn - j;
        j = jn;
        t = qn - baseN;
        load_slice(dnext);
        take(lane);
        live = true;
    }

    __device__ __forceinline__ void start(const int4 *tab_, int *counter_, const uint2 *slices_, const unsigned char *blob_, int nj_, int lane)
    {
        tab = tab_; counter = counter_; slices = slices_; blob = blob_; nj = nj_;
        j = 0; jn = 0; baseN = 0; switchesOwed = 0; live = false;
        const int4 e = __ldg(tab);
        sliceStartN = e.z;
        nslN = e.w - e.z;
        take(lane);
        next_slice(lane);
    }
};

template <int KCE>
__device__ __forceinline__ PMeta issue_pchunk(PWalker<KCE> &wk, uint32_t slotAddr, uint32_t barAddr, int lane, uint64_t streamPolicy)
{
    constexpr uint32_t kValBytes = static_cast<uint32_t>(slot_val_bytes(KCE));
    PMeta m;
    m.kc = 0; m.flags = 0; m.t = 0;
    if (wk.switchesOwed > 0) { /* the consumer changes buffers here */
        wk.switchesOwed -= 1;
        m.flags = 4 | 8;
        return m;
    }
    if (!wk.live) return m;
    m.t = wk.t;
    const unsigned char *src0, *src1;
    uint32_t b0, b1;
    m.flags = 4;
    if (wk.ci < wk.nE) {
        const int k = wk.ci * KCE;
        m.kc = min(KCE, wk.w - k);
        src0 = wk.base + static_cast<uint32_t>(k) * 512u;
        b0 = static_cast<uint32_t>(m.kc) * 512u;
        src1 = wk.base + static_cast<uint32_t>(wk.w) * 512u + static_cast<uint32_t>(k >> 2) * 512u;
        b1 = static_cast<uint32_t>((m.kc + 3) >> 2) * 512u;
    } else {
        const int k = (wk.ci - wk.nE) * KCE;
        const unsigned char *rem = wk.base + static_cast<uint32_t>(wk.w) * 512u + static_cast<uint32_t>((wk.w + 3) >> 2) * 512u;
        m.kc = max(0, min(KCE, wk.wr - k));
        m.flags |= 1;
        src0 = rem + static_cast<uint32_t>(k) * 512u;
        b0 = static_cast<uint32_t>(m.kc) * 512u;
        src1 = rem + static_cast<uint32_t>(wk.wr) * 512u + static_cast<uint32_t>(k >> 2) * 512u;
        b1 = static_cast<uint32_t>((m.kc + 3) >> 2) * 512u;
    }
    if (lane == 0 && b0) {
        mbar_expect_tx(barAddr, b0 + b1);
        if (streamPolicy) {
            tma_bulk_g2s_hint(slotAddr, src0, b0, barAddr, streamPolicy);
            tma_bulk_g2s_hint(slotAddr + kValBytes, src1, b1, barAddr, streamPolicy);
        } else {
            tma_bulk_g2s(slotAddr, src0, b0, barAddr);
            tma_bulk_g2s(slotAddr + kValBytes, src1, b1, barAddr);
        }
    }
    if (++wk.ci == wk.nc) {
        m.flags |= 2;
        wk.next_slice(lane);
    }
    return m;
}

/* Multi-GPU build, cold path of the staging: this warp's share of a partition's cache list reaches
 * into the halo columns (>= n, the tail of the ascending list).  Waits (once per product and warp)
 * for the neighbours' push, reads the halo buffer - peer-written, L2 is the point of coherence -
 * with plain stores into the cache, and lets every lane arrive on the cache barrier itself
 * (release) once its asynchronous gathers have landed.  Returns the new haloReady. */
__device__ __noinline__ bool stage_halo_columns(const MainArgs &a, const int32_t *cols, int cacheCount, double *cache, uint32_t cacheBar,
                                                bool haloReady, unsigned long long *tr)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, stride = blockDim.x;
    if (a.peer.flags != nullptr && !haloReady) {
        const unsigned long long tw = tr ? global_timer_ns() : 0ull;
        peer_wait(a.peer.flags, a.peer.recvMask, a.peer.nranks, a.peer.peerPushCtas, false, a.peer.epoch, a.peer.timeoutNs, a.peer.status);
        if (tr && lane == 0) {
            const unsigned long long te = global_timer_ns();
            atomicMax(tr + 6, te - tw);
            atomicMax(tr + 7, te);
        }
    }
    for (int i = warp * 32 + lane; i < cacheCount; i += stride) {
        const int c = __ldg(cols + i);
        if (c >= a.n) cache[i] = ld_halo_f64(a.peer, a.n, c);
    }
    cp_async_wait_all();
    mbar_arrive(cacheBar);
    return true;
}

/* NS = staging slots per warp (2 or 3): the depth of a warp's chunk pipeline.  16 warps x 3 slots with
 * the 128-register budget of a 512-thread CTA beat 20 x 2 at 96 registers wherever the two buffers
 * leave 120 KB (27-point 256^3: 795 -> see profiles/r2_notes.md); measured: the register budget
 * matters as much as the bytes in flight - the consumer's LDS chains decide how fast a slot turns around. */
/* DOT: the build that also accumulates a.dot += sum_r y[r] * x[r] (ehyb_spmv_dot, the p.Ap of a CG
 * iteration).  A build of its own, and the warp's partial sum lives in shared memory, not in a
 * register pair: the consumer loop sits at the register cap of every build (the dot folded into the
 * 96-register multi-GPU build spilled and cost 19 % of the product, ncu). */
template <int kMaxThreads, int KCE, bool PEER, int NS, bool DOT = false>
__global__ void __launch_bounds__(kMaxThreads, 1) ehyb_persistent_kernel(const __grid_constant__ MainArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr uint32_t kSlotBytes = static_cast<uint32_t>(slot_bytes(KCE));
    constexpr uint32_t kSlotValBytes = static_cast<uint32_t>(slot_val_bytes(KCE));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = blockDim.x >> 5;
    /* this CTA's work items: {rowStart, rowEnd, sliceStart, sliceEnd} of the partition - the slices
     * restricted to the CTA's run -, {cacheStart, cacheCount, flags (1: the cache list has halo
     * columns), row of the run's first slice} */
    const int item0 = __ldg(a.ctaStart + blockIdx.x);
    const int nj = __ldg(a.ctaStart + blockIdx.x + 1) - item0;
    const int4 *tab = reinterpret_cast<const int4 *>(a.ctaTab) + 2 * static_cast<size_t>(item0);

    /* header: [0,16) window bars, [16,32) cache bars, [32,48) empty bars, [64,640) slot bars,
     * [640,832) DOT builds: one partial sum per warp, [960,964) sequence counter */
    const uint32_t hdr = smem_u32(smem);
    int *seqCounter = reinterpret_cast<int *>(smem + 960);
    const uint32_t winBytes = (static_cast<uint32_t>(a.W + 2) * 8u + 127u) & ~127u;
    const uint32_t cacheBytes = (static_cast<uint32_t>(a.cacheCap) * 8u + 127u) & ~127u;
    const uint32_t bufBytes = winBytes + cacheBytes;
    unsigned char *buf0 = smem + kPersistHeader;
    const uint32_t slotBar0 = hdr + 64u + static_cast<uint32_t>(warp * NS) * 8u;
    const uint32_t slot0 = smem_u32(buf0) + 2u * bufBytes + static_cast<uint32_t>(warp * NS) * kSlotBytes;
    const bool pusher = PEER && a.peer.flags != nullptr && static_cast<int>(blockIdx.x) < a.peer.pushCtas && warp == nw - 1;
    /* development (EHYB_TRACE=1, PEER build): 8 stamps per CTA - start, previous grid complete, push
     * done, first window staged, last warp done, SM, longest wait for the neighbours' flags (ns),
     * when that wait ended */
    unsigned long long *tr = PEER && a.trace ? a.trace + static_cast<size_t>(blockIdx.x) * 8 : nullptr;
    if (PEER && tr && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tr[0] = global_timer_ns();
        tr[5] = smid;
        tr[2] = tr[4] = tr[6] = tr[7] = 0;
    }

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (nj <= 0) { /* fewer slices than CTAs: nothing but this CTA's share of the halo push */
        if (PEER && pusher) {
            const bool ok = peer_push_check(a.peer);
            asm volatile("griddepcontrol.wait;" ::: "memory");
            peer_push_all(a.peer, a.x, lane, ok);
        }
        return;
    }
    if (tid == 0) {
        *seqCounter = 0;
        for (int b = 0; b < 2; ++b) {
            mbar_init(hdr + 8u * b, 1);                              /* window: the TMA issuer's arrive.expect_tx */
            mbar_init(hdr + 16u + 8u * b, static_cast<uint32_t>(nw * 32)); /* cache: every lane arrives (through cp.async) */
            mbar_init(hdr + 32u + 8u * b, static_cast<uint32_t>(nw)); /* empty: one arrival per warp */
        }
        for (int i = 0; i < nw * NS; ++i) mbar_init(hdr + 64u + i * 8u, 1);
        fence_mbar_init();
    }
    __syncthreads();

    const uint64_t streamPolicy = a.l2hint ? make_evict_first_policy() : 0ull;
    uint64_t keepPolicy; /* x window: evict-last with the L2 hints, normal priority without */
    if (a.l2hint) keepPolicy = make_evict_last_policy();
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(keepPolicy));
    PWalker<KCE> wk;
    wk.start(tab, seqCounter, reinterpret_cast<const uint2 *>(a.slices), a.blob, nj, lane);
    PMeta meta0, meta1, meta2; /* chunk in slot 0, 1, 2 (named, not an array: the slot index is dynamic) */
    meta0 = issue_pchunk(wk, slot0, slotBar0, lane, streamPolicy);
    meta1 = issue_pchunk(wk, slot0 + kSlotBytes, slotBar0 + 8u, lane, streamPolicy);
    meta2.kc = 0; meta2.flags = 0; meta2.t = 0;
    if (NS > 2) meta2 = issue_pchunk(wk, slot0 + 2u * kSlotBytes, slotBar0 + 16u, lane, streamPolicy);
    uint32_t phases = 0;

    /* multi-GPU, pushing warp: everything of the halo push that does not need x */
    bool pushOk = false;
    if (PEER && pusher) pushOk = peer_push_check(a.peer);

    asm volatile("griddepcontrol.wait;" ::: "memory"); /* x and y belong to the stream's previous work */
    const bool tma_ok = (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
    if (PEER && tr && tid == 0) tr[1] = global_timer_ns();
    /* This CTA's share of the x entries the neighbours need (last warp only) goes out right after
     * the warp has issued its share of the first partition's staging - unless that partition has
     * halo columns itself (table flag; the session orders those last whenever it can): a warp must
     * never wait for its neighbours' flags before it has pushed, two GPUs doing that would deadlock. */
    bool pushPending = PEER && pusher;
    if (PEER && pushPending && (__ldg(tab + 1).z & 1)) {
        peer_push_all(a.peer, a.x, lane, pushOk);
        pushPending = false;
    }

    /* Staging of partition dutyJ of this CTA into buffer dutyJ&1: its window (TMA, warp 0) and this
     * warp's share of its remainder cache.  Waits until every warp has left partition dutyJ-2
     * (`empty` barrier), fetches the warp's list entries four groups at a time and gathers x with
     * asynchronous 8-byte copies; every lane then arrives on the cache barrier through cp.async
     * (the arrival completes when its gathers have landed), so the warp does not wait for them.
     * Inlined at two places only (start-up and one site in the loop): the code size of this kernel
     * matters (a polled, non-blocking variant with checks in every iteration measured slower). */
    int dutyState = 0, dutyJ = 0; /* dutyState 1: partition dutyJ still has to be staged by this warp */
    int nxtPs = 0, nxtPe = 0, nxtRow0 = 0; /* partition staged last = the one the consumer enters next: rows, row of the item's first slice */
    bool haloReady = false;       /* PEER: this warp has seen the neighbours' flags of this product */
    auto duty_begin = [&](int jj) { dutyState = 1; dutyJ = jj; };
    auto duty_run = [&]() {
        const int b = dutyJ & 1;
        const int4 e0 = __ldg(tab + 2 * dutyJ);
        const int4 e1 = __ldg(tab + 2 * dutyJ + 1);
        const int cacheStart = e1.x, cacheCount = e1.y;
        nxtPs = e0.x; nxtPe = e0.y; nxtRow0 = e1.w;
        if (dutyJ >= 2) mbar_wait_bounded(hdr + 32u + 8u * b, static_cast<uint32_t>(((dutyJ - 2) >> 1) & 1));
        if (warp == 0) {
            const int ps_ = e0.x;
            double *win = reinterpret_cast<double *>(buf0 + static_cast<size_t>(b) * bufBytes);
            const int g0 = ps_ & ~1;
            const int len = max(0, min(ps_ + a.W, a.n) - g0);
            if (tma_ok) {
                if (lane == 0) {
                    if (len & 1) win[len - 1] = a.x[g0 + len - 1]; /* odd tail element, released by the arrive below */
                    const uint32_t bulkBytes = static_cast<uint32_t>(len & ~1) * 8u;
                    mbar_expect_tx(hdr + 8u * b, bulkBytes);
                    const char *src = reinterpret_cast<const char *>(a.x + g0);
                    const uint32_t dst = smem_u32(win);
                    for (uint32_t off = 0; off < bulkBytes; off += 32768u) tma_bulk_g2s_hint(dst + off, src + off, min(32768u, bulkBytes - off), hdr + 8u * b, keepPolicy);
                }
            } else { /* x not 16-byte aligned: plain copies by warp 0 */
                for (int i = lane; i < len; i += 32) win[i] = a.x[g0 + i];
                __syncwarp();
                if (lane == 0) mbar_arrive(hdr + 8u * b);
            }
        }
        const int32_t *cols = a.cacheCols + cacheStart;
        const uint32_t cacheA = smem_u32(buf0) + static_cast<uint32_t>(b) * bufBytes + winBytes;
        const int stride = nw * 32;
        bool halo = false; /* PEER: this lane met a halo column (>= n): the tail of the ascending list */
        for (int i0 = warp * 32 + lane; i0 < cacheCount; i0 += 4 * stride) {
            /* this warp's entries: groups of 32 dealt round-robin over the warps, 4 groups per round */
            const int c0 = __ldg(cols + i0);
            const int c1 = i0 + stride < cacheCount ? __ldg(cols + i0 + stride) : -1;
            const int c2 = i0 + 2 * stride < cacheCount ? __ldg(cols + i0 + 2 * stride) : -1;
            const int c3 = i0 + 3 * stride < cacheCount ? __ldg(cols + i0 + 3 * stride) : -1;
            if (PEER) {
                halo = halo || c0 >= a.n || c1 >= a.n || c2 >= a.n || c3 >= a.n;
                if (c0 < a.n) cp_async_8(cacheA + static_cast<uint32_t>(i0) * 8u, a.x + c0);
                if (c1 >= 0 && c1 < a.n) cp_async_8(cacheA + static_cast<uint32_t>(i0 + stride) * 8u, a.x + c1);
                if (c2 >= 0 && c2 < a.n) cp_async_8(cacheA + static_cast<uint32_t>(i0 + 2 * stride) * 8u, a.x + c2);
                if (c3 >= 0 && c3 < a.n) cp_async_8(cacheA + static_cast<uint32_t>(i0 + 3 * stride) * 8u, a.x + c3);
            } else {
                cp_async_8(cacheA + static_cast<uint32_t>(i0) * 8u, a.x + c0);
                if (c1 >= 0) cp_async_8(cacheA + static_cast<uint32_t>(i0 + stride) * 8u, a.x + c1);
                if (c2 >= 0) cp_async_8(cacheA + static_cast<uint32_t>(i0 + 2 * stride) * 8u, a.x + c2);
                if (c3 >= 0) cp_async_8(cacheA + static_cast<uint32_t>(i0 + 3 * stride) * 8u, a.x + c3);
            }
        }
        if (PEER && __any_sync(0xffffffffu, halo)) {
            /* (out of line: the code size of the loop around this staging is measurable) */
            haloReady = stage_halo_columns(a, cols, cacheCount, reinterpret_cast<double *>(buf0 + static_cast<size_t>(b) * bufBytes + winBytes),
                                           hdr + 16u + 8u * b, haloReady, tr);
        } else {
            cp_async_mbar_arrive_noinc(hdr + 16u + 8u * b); /* every lane: arrives when its gathers have landed */
        }
        dutyState = 0;
    };

    /* consumer state for partition jC */
    int jC = 0;
    bool cacheReady = false;
    duty_begin(0);
    duty_run();
    if (PEER && pushPending) peer_push_all(a.peer, a.x, lane, pushOk);
    if (PEER && tr && pusher && lane == 0) tr[2] = global_timer_ns();
    int ps = nxtPs, pe = nxtPe, row0 = nxtRow0;
    uint32_t xsAddr = smem_u32(buf0) + static_cast<uint32_t>(ps - (ps & ~1)) * 8u;
    uint32_t cacheAddr = smem_u32(buf0) + winBytes;
    if (nj > 1) duty_begin(1);
    mbar_wait_bounded(hdr + 0u, 0);
    if (PEER && tr && tid == 0) tr[3] = global_timer_ns();
    bool dutyDue = false; /* set at the end of a slice: stage the next partition at the top of the next iteration */

    double acc0 = 0.0, acc1 = 0.0, r0 = 0.0, r1 = 0.0;
    double *dotAcc = reinterpret_cast<double *>(smem + 640) + warp; /* DOT: this warp's share of y.x (header bytes [640, 832)) */
    if (DOT && lane == 0) *dotAcc = 0.0;
    int s = 0;
#pragma unroll 1
    for (;;) {
        const PMeta m = s == 0 ? meta0 : (s == 1 || NS == 2) ? meta1 : meta2;
        if (!(m.flags & 4)) break;
        const uint32_t slot = slot0 + static_cast<uint32_t>(s) * kSlotBytes;
        const uint32_t bar = slotBar0 + static_cast<uint32_t>(s) * 8u;
        /* the staging of the next partition, at its single site in the loop: after this warp's
         * first slice of the current partition, at the latest before it leaves the partition */
        if (dutyState != 0 && (dutyDue || (m.flags & 8))) duty_run();
        dutyDue = false;
        if (m.flags & 8) {
            /* this warp is done with partition jC: tell the others, move to the other buffer */
            __syncwarp();
            if (lane == 0) mbar_arrive(hdr + 32u + 8u * static_cast<uint32_t>(jC & 1));
            jC += 1;
            const uint32_t b = static_cast<uint32_t>(jC & 1), par = static_cast<uint32_t>((jC >> 1) & 1);
            ps = nxtPs; pe = nxtPe; row0 = nxtRow0; /* (the staging of partition jC by this warp came last) */
            xsAddr = smem_u32(buf0) + b * bufBytes + static_cast<uint32_t>(ps - (ps & ~1)) * 8u;
            cacheAddr = smem_u32(buf0) + b * bufBytes + winBytes;
            cacheReady = false;
            if (jC + 1 < nj) duty_begin(jC + 1);
            mbar_wait_bounded(hdr + 8u * b, par);
        } else {
            if (m.kc) {
                mbar_wait_bounded(bar, (phases >> s) & 1u);
                phases ^= 1u << s;
                if ((m.flags & 1) && !cacheReady) {
                    mbar_wait_bounded(hdr + 16u + 8u * static_cast<uint32_t>(jC & 1), static_cast<uint32_t>((jC >> 1) & 1));
                    cacheReady = true;
                }
                const uint32_t vAddr = slot + static_cast<uint32_t>(lane) * 16u;
                if (!(a.dbg & ((m.flags & 1) ? 1 : 2))) {
                    const bool rem = (m.flags & 1) != 0;
                    const uint32_t xb = rem ? cacheAddr : xsAddr;
                    double s0 = rem ? r0 : acc0, s1 = rem ? r1 : acc1;
                    const uint32_t cAddr = slot + kSlotValBytes + static_cast<uint32_t>(lane) * 16u;
                    const int nfull = m.kc >> 2;
#pragma unroll 2
                    for (int g = 0; g < nfull; ++g) {
                        const uint4 c = lds_u32x4(cAddr + g * 512u);
                        const double2 v0 = lds_f64x2(vAddr + (4 * g + 0) * 512u);
                        const double2 v1 = lds_f64x2(vAddr + (4 * g + 1) * 512u);
                        const double2 v2 = lds_f64x2(vAddr + (4 * g + 2) * 512u);
                        const double2 v3 = lds_f64x2(vAddr + (4 * g + 3) * 512u);
                        const double x00 = lds_f64(xaddr_lo(c.x, xb)), x01 = lds_f64(xaddr_lo(c.z, xb));
                        const double x10 = lds_f64(xaddr_hi(c.x, xb)), x11 = lds_f64(xaddr_hi(c.z, xb));
                        const double x20 = lds_f64(xaddr_lo(c.y, xb)), x21 = lds_f64(xaddr_lo(c.w, xb));
                        const double x30 = lds_f64(xaddr_hi(c.y, xb)), x31 = lds_f64(xaddr_hi(c.w, xb));
                        s0 = fma(v0.x, x00, s0);
                        s1 = fma(v0.y, x01, s1);
                        s0 = fma(v1.x, x10, s0);
                        s1 = fma(v1.y, x11, s1);
                        s0 = fma(v2.x, x20, s0);
                        s1 = fma(v2.y, x21, s1);
                        s0 = fma(v3.x, x30, s0);
                        s1 = fma(v3.y, x31, s1);
                    }
                    const int tail = m.kc & 3;
                    if (tail) {
                        const uint4 c = lds_u32x4(cAddr + nfull * 512u);
                        const uint32_t cols0[3] = {c.x & 0xffffu, c.x >> 16, c.y & 0xffffu};
                        const uint32_t cols1[3] = {c.z & 0xffffu, c.z >> 16, c.w & 0xffffu};
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            if (i < tail) {
                                const double2 v = lds_f64x2(vAddr + (4 * nfull + i) * 512u);
                                s0 = fma(v.x, lds_f64(xb + cols0[i] * 8u), s0);
                                s1 = fma(v.y, lds_f64(xb + cols1[i] * 8u), s1);
                            }
                        }
                    }
                    if (rem) { r0 = s0; r1 = s1; } else { acc0 = s0; acc1 = s1; }
                }
            }
            if (m.flags & 2) {
                const int r = row0 + m.t * EHYB_SLICE_ROWS + lane;
                if (r < pe) a.y[r] = acc0 + r0;
                if (r + 32 < pe) a.y[r + 32] = acc1 + r1;
                if constexpr (DOT) {
                    double v = dot_rows(a.x, xsAddr, ps, min(ps + a.W, a.n), pe, r, acc0 + r0, acc1 + r1, 0.0);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    if (lane == 0) *dotAcc += v;
                }
                acc0 = acc1 = r0 = r1 = 0.0;
                dutyDue = true;
            }
        }
        __syncwarp();
        const PMeta mn = issue_pchunk(wk, slot, bar, lane, streamPolicy);
        if (s == 0) meta0 = mn; else if (s == 1 || NS == 2) meta1 = mn; else meta2 = mn;
        s = s + 1 == NS ? 0 : s + 1;
    }
    if constexpr (DOT) {
        if (lane == 0) atomicAdd(a.dot, *dotAcc);
    }
    if (PEER && tr && lane == 0) atomicMax(tr + 4, global_timer_ns());
    /* (every warp has passed all nj-1 switch markers here: the walker always ends in the last
     * partition, and a marker completes the staging it owes before it leaves a partition) */
}

/* ---------------------------------------------------------------- overflow kernel -- */

/*
 * COO remainder (row-sorted): each warp takes perWarp consecutive entries, 32 at a time (the
 * launcher keeps perWarp at 32 for short lists, so that the kernel is one wave of independent
 * warps instead of a few warps walking dependent loads; long lists use up to kOvfPerWarp);
 * products are reduced per row with a warp-shuffle segmented scan and the last lane of every
 * row segment adds its sum to y (the row's ELL + in-slice part is already there: this kernel
 * runs after ehyb_main_kernel on the same stream).  A segment that continues into the next
 * group of 32 is carried in registers instead of being flushed.
 */
constexpr int kOvfPerWarp = 32 * 8;
/* kOvfUnroll = groups of 32 entries whose loads and gathers are in flight together */
template <int kOvfUnroll>
__global__ void __launch_bounds__(256) ehyb_overflow_kernel(const OverflowArgs a)
{
    /* This grid is launched programmatically behind the main kernel and waits here for its y.
     * The next product's main kernel is released only when this CTA is done (trigger at the end,
     * a.lateTrigger): released at the top, its 768-thread CTAs take over the SMs while most of
     * this grid's CTAs have not been scheduled yet and then sit waiting for them (measured:
     * 211 us per product instead of 144 on R-MAT scale 20). */
    if (!a.lateTrigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int lane = threadIdx.x & 31;
    const int64_t warpId = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t begin = warpId * a.perWarp;
    if (begin >= a.count) return; /* (an exited CTA counts as triggered) */
    const int64_t end = min(begin + a.perWarp, a.count);
    int carryRow = -1;
    double carry = 0.0;
    bool haloReady = false;
    for (int64_t base = begin; base < end; base += 32 * kOvfUnroll) {
        /* all the streaming loads of kOvfUnroll groups first, then all the x gathers: a warp keeps
         * 4 x 32 independent gathers in flight instead of 32 (the list is L2-latency bound) */
        int r[kOvfUnroll], c[kOvfUnroll];
        double v[kOvfUnroll], xv[kOvfUnroll];
        bool anyHalo = false;
#pragma unroll
        for (int u = 0; u < kOvfUnroll; ++u) {
            const int64_t i = base + 32 * u + lane;
            const bool live = i < end;
            r[u] = live ? __ldg(a.row + i) : -2 - lane; /* dead lanes: unique rows, never merged, never written */
            c[u] = live ? __ldg(a.col + i) : 0;
            v[u] = live ? __ldg(a.val + i) : 0.0;
            anyHalo = anyHalo || c[u] >= a.n;
        }
        if (a.peer.flags != nullptr && !haloReady && __any_sync(0xffffffffu, anyHalo)) {
            /* first halo column of this warp: wait (once) for the neighbours' push */
            peer_wait(a.peer.flags, a.peer.recvMask, a.peer.nranks, a.peer.peerPushCtas, false, a.peer.epoch, a.peer.timeoutNs, a.peer.status);
            haloReady = true;
        }
#pragma unroll
        for (int u = 0; u < kOvfUnroll; ++u) xv[u] = c[u] < a.n ? ld_gather_f64(a.x + c[u]) : ld_halo_f64(a.peer, a.n, c[u]);
#pragma unroll
        for (int u = 0; u < kOvfUnroll; ++u) {
            const int64_t gbase = base + 32 * u;
            if (gbase >= end) break; /* warp-uniform */
            const bool live = gbase + lane < end;
            const int row = r[u];
            double prod = v[u] * xv[u];
            if (lane == 0 && row == carryRow) prod += carry; /* continue the carried segment */
            const int flushRow = (lane == 0 && carryRow >= 0 && row != carryRow) ? carryRow : -1;
            if (flushRow >= 0) atomicAdd(a.y + flushRow, carry);
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, prod, off);
                const int rr = __shfl_up_sync(0xffffffffu, row, off);
                if (lane >= off && rr == row) prod += t;
            }
            const int rnext = __shfl_down_sync(0xffffffffu, row, 1);
            const bool tailOfSeg = live && (lane == 31 || rnext != row);
            /* the segment ending in lane 31 may continue in the next group: carry it */
            const bool carries = lane == 31 && live && gbase + 32 < end;
            if (tailOfSeg && !carries) atomicAdd(a.y + row, prod);
            carryRow = __shfl_sync(0xffffffffu, carries ? row : -1, 31);
            carry = __shfl_sync(0xffffffffu, prod, 31);
        }
    }
    if (lane == 0 && carryRow >= 0) atomicAdd(a.y + carryRow, carry);
    if (a.lateTrigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

/* ---------------------------------------------------------------- overflow stream --- */

/*
 * Large overflow lists (power-law graphs: the list is the matrix) in the tile-packed CSR-like
 * stream of host/ovfstream.c: 12.4 bytes per entry, hub columns in shared memory, no atomics.
 * One persistent CTA per SM; a warp takes tiles of TG x 32 consecutive entries (tile t -> warp
 * t mod #warps: neighbouring warps stream neighbouring tiles).
 *
 *   - a tile is ONE record (values, columns, group words, flags) that lane 0 moves into one of the
 *     warp's staging slots with a bulk copy (UBLKCP, evict-first) `slots` tiles ahead: the matrix
 *     stream costs no registers and no exposed latency.  The only global loads a warp waits for are
 *     its x gathers and - issued WITH them - the rows of the segments that end in each lane: one
 *     memory latency per tile, where plain loads of three arrays walked three dependent ones
 *     (stream -> gather -> row of the segment; R-MAT 24: 2 278 us, behind the COO list);
 *   - the x values of the hub columns (the most referenced ones) are gathered into shared memory
 *     once per CTA; an entry with the top bit of its column set reads there (LDS) instead of
 *     occupying one 32-byte L1TEX sector per lane;
 *   - row segments: lane j of a group belongs to segment seg0 + popc(mask & bits 1..j); products
 *     are summed per segment with a warp-shuffle segmented scan (a lane knows its distance to the
 *     head of its segment from the mask: only the sums travel), carried from group to group inside
 *     the tile in registers;
 *   - a segment that lies inside the tile is stored by the lane at its end: y[row] = sum (or +=
 *     when the main kernel has written the row's slice part);
 *   - the tile's first segment if it continues a row of the previous tile, and its last one if the
 *     row goes on in the next tile, go to the tile's two CARRY slots instead (their rows are fixed
 *     by the data and were written at upload); ehyb_ovfstream_fixup adds the slots of a row in tile
 *     order.  Who adds what, and in which order, is fixed by the data: y is bit-reproducible (the
 *     COO kernel's atomics are not).
 */
struct OvfStreamArgs {
    const unsigned char *tiles; /* nTiles records of EHYB_OVF_TILE_BYTES(TG) */
    const int32_t *rowOfSeg;
    const int32_t *hubCols;
    int nHub;
    int nTiles;
    int slots;                 /* staging slots per warp (2..4) */
    const double *x;
    double *y;
    double *carryVal;          /* [2 * nTiles]: head and tail slot of every tile */
    int accumulate;            /* 1: y[row] += sum (the main kernel wrote y), 0: y[row] = sum (y was zeroed) */
};

constexpr int kStreamHeader = 1024; /* mbarriers: warps x slots x 8 bytes */
constexpr int kStreamMaxSlots = 4;

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

template <int TG, int kThreads>
__global__ void __launch_bounds__(kThreads, 1) ehyb_ovfstream_kernel(const __grid_constant__ OvfStreamArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr uint32_t E = 32u * TG;
    constexpr uint32_t kTileBytes = EHYB_OVF_TILE_BYTES(TG);
    constexpr uint32_t kColOff = 8u * E, kGrpOff = 12u * E, kFlagOff = 12u * E + 8u * TG;
    constexpr int kWarps = kThreads / 32;
    static_assert(kWarps * kStreamMaxSlots * 8 <= kStreamHeader, "barrier header too small");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slots = a.slots;
    const uint32_t bar0 = smem_u32(smem) + static_cast<uint32_t>(warp * slots) * 8u;
    const uint32_t slot0 = smem_u32(smem) + kStreamHeader + static_cast<uint32_t>(warp * slots) * kTileBytes;
    double *hub = reinterpret_cast<double *>(smem + kStreamHeader + static_cast<size_t>(kWarps * slots) * kTileBytes);
    if (tid == 0) {
        for (int i = 0; i < kWarps * slots; ++i) mbar_init(smem_u32(smem) + 8u * i, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const uint64_t pol = make_evict_first_policy();
    const int Wt = static_cast<int>(gridDim.x) * kWarps;
    const int first = static_cast<int>(blockIdx.x) * kWarps + warp;
    /* the first `slots` tiles of this warp: constant data, on their way while the hubs are gathered */
    if (lane == 0) {
        for (int q = 0; q < slots; ++q) {
            const int64_t t = static_cast<int64_t>(first) + static_cast<int64_t>(q) * Wt;
            if (t < a.nTiles) {
                mbar_expect_tx(bar0 + 8u * q, kTileBytes);
                tma_bulk_g2s_hint(slot0 + static_cast<uint32_t>(q) * kTileBytes, a.tiles + static_cast<size_t>(t) * kTileBytes, kTileBytes, bar0 + 8u * q, pol);
            }
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory"); /* x, and y as the main kernel / the memset left it */
    for (int i = tid; i < a.nHub; i += kThreads) hub[i] = ld_gather_f64(a.x + __ldg(a.hubCols + i));
    __syncthreads();
    const uint32_t hubAddr = smem_u32(hub);
    const uint32_t lanemaskLe = (2u << lane) - 1u; /* bits 0..lane */
    uint32_t phases = 0;
    int s = 0;
#pragma unroll 1
    for (int64_t tile = first; tile < a.nTiles; tile += Wt) {
        const uint32_t slot = slot0 + static_cast<uint32_t>(s) * kTileBytes, bar = bar0 + static_cast<uint32_t>(s) * 8u;
        mbar_wait_bounded(bar, (phases >> s) & 1u);
        phases ^= 1u << s;
        /* Lane j owns the FOUR CONSECUTIVE entries 4j..4j+3 of the tile: it adds what belongs to one row
         * serially, and only one value per lane - the sum behind its last row start - goes through a
         * cross-lane segmented scan: 12 shuffles per tile.  (One entry per lane and a scan per group of
         * 32 was 60: ncu showed 38 % of the LSU data pipe - the pipe this kernel is bound by - busy
         * with shared-memory-class wavefronts of which the loads were 8 %: the shuffles.) */
        static_assert(TG == 4, "four entries per lane");
        const uint4 cc = lds_u32x4(slot + kColOff + static_cast<uint32_t>(lane) * 16u);
        const int2 gmA = lds_s32x2(slot + kGrpOff + static_cast<uint32_t>(lane >> 3) * 8u); /* this lane's group of 32: {seg0, mask} */
        const int2 gm3 = lds_s32x2(slot + kGrpOff + 24u);
        const int headSeg = static_cast<int>(lds_u32(slot + kGrpOff));
        const uint32_t flags = lds_u32(slot + kFlagOff);
        const uint32_t gmask = static_cast<uint32_t>(gmA.y);
        const int sh = 4 * (lane & 7);
        const uint32_t nib = (gmask >> sh) & 0xfu;                /* bit k: entry 4 lane + k starts a new row */
        const int nStart = __popc(nib);
        /* segment of the lane's first entry, and the segment BEFORE the lane's first row start (= the one
         * the entries in front of it, and the lanes before, contribute to) */
        const int seg0 = gmA.x + __popc(gmask & ((2u << sh) - 1u) & ~1u);
        const int sB = seg0 - static_cast<int>(nib & 1u);
        const int tailSeg = gm3.x + __popc(static_cast<uint32_t>(gm3.y) & ~1u);
        const bool preStore = nStart > 0 && !(lane == 0 && (nib & 1u)); /* a segment ends at this lane's first row start */
        /* the gathers, and with them the rows of the segments this lane will store: all in flight together */
        double p[4];
        {
            const uint32_t cs[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                p[k] = (cs[k] & EHYB_OVF_HUB_BIT) ? lds_f64(hubAddr + (cs[k] & 0x7fffffffu) * 8u) : __ldg(a.x + cs[k]); /* (plain launch: x is constant while this grid lives) */
        }
        const int rowPre = preStore ? __ldg(a.rowOfSeg + sB) : -1;
        int rowIn[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) rowIn[k] = nStart > k + 1 ? __ldg(a.rowOfSeg + sB + 1 + k) : -1; /* complete segments inside the lane */
        const int rowTail = lane == 31 ? __ldg(a.rowOfSeg + tailSeg) : -1;
        {
            const double2 v01 = lds_f64x2(slot + static_cast<uint32_t>(lane) * 32u);
            const double2 v23 = lds_f64x2(slot + static_cast<uint32_t>(lane) * 32u + 16u);
            p[0] *= v01.x; p[1] *= v01.y; p[2] *= v23.x; p[3] *= v23.y;
        }
        /* Refill the slot - but only when every lane's reads of it have been PERFORMED, not merely
         * issued: this kernel keeps the LSU pipe busy with 32-wavefront gathers, a shared-memory load
         * can wait in its queue for microseconds, and the bulk copy of the next tile (async proxy) would
         * then land under it.  Measured: without the fence a few rows of R-MAT 24 left the gate, more
         * with more slots (profiles/r2_notes.md).  fence.proxy.async orders each lane's generic-proxy
         * reads before later async-proxy writes; the barrier extends that to the lane that issues. */
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const int64_t t = tile + static_cast<int64_t>(slots) * Wt;
            if (t < a.nTiles) {
                mbar_expect_tx(bar, kTileBytes);
                tma_bulk_g2s_hint(slot, a.tiles + static_cast<size_t>(t) * kTileBytes, kTileBytes, bar, pol);
            }
        }
        const bool headCont = (flags & 1u) != 0u, tailCont = (flags & 2u) != 0u;
        auto store = [&](int sg, int r, double sum) {
            if (r < 0) return; /* the padding entries behind the end of the list */
            const bool isHead = headCont && sg == headSeg, isTail = tailCont && sg == tailSeg;
            if (isHead) a.carryVal[2 * tile] = sum;               /* (also the tail: that slot keeps its 0) */
            else if (isTail) a.carryVal[2 * tile + 1] = sum;
            else if (a.accumulate) atomicAdd(a.y + r, sum); /* RED.ADD: nobody else adds to this row in this launch, and the warp does not wait for y */
            else a.y[r] = sum;
        };
        /* inside the lane: `head` = what precedes the first row start, complete segments between two row
         * starts are stored on the spot, `run` ends as what follows the last row start */
        double run = 0.0, head = 0.0;
        int seen = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if ((nib >> k) & 1u) {
                if (seen == 0) head = run;
                else store(sB + seen, seen == 1 ? rowIn[0] : seen == 2 ? rowIn[1] : rowIn[2], run); /* (no dynamic index: registers) */
                run = 0.0;
                ++seen;
            }
            run += p[k];
        }
        /* across the lanes: inclusive scan of `run`, restarted at every lane that has a row start */
        const uint32_t startLanes = __ballot_sync(0xffffffffu, nStart > 0);
        const uint32_t upto = startLanes & lanemaskLe;
        const int dist = upto ? lane - (31 - __clz(static_cast<int>(upto))) : lane; /* lanes back to the last restart (or to lane 0) */
        double sc = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, sc, off);
            if (dist >= off) sc += t;
        }
        double before = __shfl_up_sync(0xffffffffu, sc, 1); /* what the lanes before add to the segment that ends here */
        if (lane == 0) before = 0.0;
        if (preStore) store(sB, rowPre, before + head);
        if (lane == 31) store(tailSeg, rowTail, sc);               /* the tile's last segment */
        s = s + 1 == slots ? 0 : s + 1;
    }
}

/* Adds the carry slots of every row that spans tiles.  The slots of a row are a run of consecutive
 * slots known when the stream is built (host/ovfstream.c: runs = {first slot, slots, row}, the short
 * runs first): one THREAD per short run, one WARP per long one - lane l adds slots l, l + 32, ... in
 * order and the lanes are combined by a fixed shuffle tree, so the sum of a row that spans thousands
 * of tiles (a hub row of a power-law graph) is still fixed by the data and no longer one thread's
 * serial chain of dependent loads (R-MAT 24: ~500 us of a 2 000 us product before). */
__global__ void __launch_bounds__(256) ehyb_ovfstream_fixup(const int32_t *__restrict__ runs, int64_t nShort, int64_t nRuns,
                                                          const double *__restrict__ carryVal, double *__restrict__ y, int accumulate)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int64_t shortBlocks = (nShort + 255) / 256;
    if (static_cast<int64_t>(blockIdx.x) < shortBlocks) {
        const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
        if (i >= nShort) return;
        const int first = runs[3 * i], len = runs[3 * i + 1], r = runs[3 * i + 2];
        double sum = 0.0;
        for (int k = 0; k < len; ++k) sum += carryVal[first + k];
        y[r] = accumulate ? y[r] + sum : sum;
        return;
    }
    const int lane = threadIdx.x & 31;
    const int64_t i = nShort + (static_cast<int64_t>(blockIdx.x) - shortBlocks) * 8 + (threadIdx.x >> 5);
    if (i >= nRuns) return;
    const int first = runs[3 * i], len = runs[3 * i + 1], r = runs[3 * i + 2];
    double sum = 0.0;
    for (int k = lane; k < len; k += 32) sum += carryVal[first + k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (lane == 0) y[r] = accumulate ? y[r] + sum : sum;
}

} /* namespace ehyb */

