/*
 * ehyb_device.cu -- device management behind the C ABI (include/ehyb.h, "device session").
 *
 * Replaces the reference's cudaMallocTransDataEHYB (spmv.cu:6-60: 12 cudaMalloc + 11 blocking
 * copies per session, sizes computed with the wrong element type, B-5) and the per-call
 * launcher matrixVectorBlockELL (kernel.cu:324-380: cudaFuncSetAttribute on every product,
 * three launches, legacy default stream) with:
 *   - one upload per array into persistent buffers, on the session's own stream;
 *   - shared-memory opt-in set once; the product is one launch (two when the overflow list
 *     is not empty), optionally replayed from a CUDA graph;
 *   - an L2 access-policy window (persisting) on x, so that the remainder gathers keep
 *     hitting L2 while the matrix streams through with evict-first loads;
 *   - CUDA-event timing on the session stream;
 *   - int status codes, no exit().
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../host/common.h"
#include "../host/ovfstream.h"
#include "ehyb_kernels.cuh"

using namespace ehyb;

/* the chunk size of the staged kernel is compile-time (see ChunkWalker); the variants below are
 * selectable with EHYB_CHUNK for experiments, the default is 4 columns (2.5 KB slots) */

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return ehyb_fail(EHYB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

constexpr int kMaxOvfBlocks = 64;

struct ehyb_handle {
    int device;
    cudaStream_t stream, h2d, d2h;
    int64_t n, ncols, nnz, nOvf, blobBytes, algBytes;
    int nParts, W, kpp, nSlices, threads, ctasPerSM, kernel, kcEll, kcRem, grid, slotsPerWarp;
    size_t smemBytes;
    ehyb_part_desc *parts;
    ehyb_slice_desc *slices;
    unsigned char *blob;
    int32_t *ovfRow, *ovfCol;
    double *ovfVal;
    int32_t *cacheCols;
    int32_t *order;      /* CTA slot -> partition (NULL: identity), see ehyb_staged_kernel */
    int32_t *ctaTab;     /* persistent kernel: work items in CTA order (8 ints each), see build_cta_tab */
    int32_t *ctaStart;   /* persistent kernel: [grid + 1] first item of every CTA */
    unsigned long long *trace; /* development: per-CTA timeline of the last product (EHYB_TRACE=1) */
    int cacheCap;        /* elements of the shared-memory remainder cache */
    int smCount;
    double *x, *y;       /* session vectors */
    double *xb[2], *yb[2]; /* double buffers of the pipelined host path (lazy) */
    cudaEvent_t evX[2], evK[2], evY[2], ev0, ev1;
    int use_graph, l2_persist, pdl, dbgSkip, haloInOverflow, l2hint, winPiece, dynamicDeal, ovfUnroll, ovfLateTrigger, pdlOvf, skipMain, prologueBarrier;
    cudaGraphExec_t gexec;
    const double *gx;
    double *gy;
    /* multi-GPU, peer-memory exchange: word in mapped pinned host memory that a kernel sets when a
     * wait on a neighbour ran into the time limit (sticky); NULL for single-GPU sessions */
    volatile uint32_t *peerStatus_h;
    /* large overflow lists: the CSR-like stream format (host/ovfstream.c) instead of the COO list */
    int ovfStream, ovfHubs, ovfTileGroups, ovfSlots, ovfBlocks; /* ovfHubs: the most any block has */
    struct OvfBlock { /* one column block = one stream, launched in block order */
        unsigned char *tiles;
        int32_t *rowOfSeg, *hubCols, *runs;
        double *carryVal;
        int nHub, nTiles;
        int64_t nRuns, nRunsShort;
    } ovfBlock[kMaxOvfBlocks];
    int64_t ovfDeviceBytes, ovfHubRefs;
    int forcePeerBuild; /* development ($EHYB_FORCE_PEER_BUILD): single-GPU sessions run the multi-GPU build of the kernel */
};

/* after a synchronisation: a product whose halo never arrived must not pass for a result */
static int peer_check(const ehyb_handle *h)
{
    if (h->peerStatus_h && *h->peerStatus_h)
        return ehyb_fail(EHYB_ERR_PEER, "a neighbour GPU did not deliver (or release) its halo within the time limit "
                                        "($EHYB_P2P_TIMEOUT_MS): the products since then are not valid");
    return EHYB_OK;
}

extern "C" int ehyb_device_count(int *count)
{
    if (!count) return ehyb_fail(EHYB_ERR_ARG, "ehyb_device_count: NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return ehyb_fail(EHYB_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return EHYB_OK;
}

extern "C" int ehyb_device_query(int device, ehyb_device_info *out)
{
    if (!out) return ehyb_fail(EHYB_ERR_ARG, "ehyb_device_query: NULL");
    cudaDeviceProp p;
    CU(cudaGetDeviceProperties(&p, device));
    memset(out, 0, sizeof *out);
    out->device = device;
    out->sm_count = p.multiProcessorCount;
    out->smem_optin_bytes = (int)p.sharedMemPerBlockOptin;
    out->smem_per_sm_bytes = (int)p.sharedMemPerMultiprocessor;
    out->l2_bytes = p.l2CacheSize;
    out->cc_major = p.major;
    out->cc_minor = p.minor;
    out->max_persist_l2_bytes = p.persistingL2CacheMaxSize;
    out->hbm_bytes = p.totalGlobalMem;
    snprintf(out->name, sizeof out->name, "%.63s", p.name);
    return EHYB_OK;
}

extern "C" void ehyb_session_opts_default(ehyb_session_opts *o)
{
    memset(o, 0, sizeof *o);
    o->device = 0;
    o->threads = 0;
    o->use_graph = 1;
    o->l2_persist_x = 1;
    o->halo_cols = 0;
    o->kernel = 0;
}

static int env_int_early(const char *name, int dflt)
{
    const char *s = getenv(name);
    return s && s[0] ? atoi(s) : dflt;
}

/* kernel variants: register budget follows the number of resident threads per SM */
typedef void (*main_kernel_t)(const MainArgs);
static main_kernel_t pick_kernel(int kernel, int threads, int ctasPerSM)
{
    if (kernel == EHYB_KERNEL_STAGED) return NULL; /* see pick_staged */
    /* 64 K registers per SM: <= 64 per thread at 1024 resident threads, <= 32 at 2048 */
    if (threads * ctasPerSM > 1024) return ehyb_main_kernel<1024, 2>;
    return ehyb_main_kernel<1024, 1>;
}

/* staged kernel builds: 4-column chunks (2.5 KB slots: most warps for the staging capacity; 8 and
 * 16 columns measured slower), x2 register budgets, x2 with / without the multi-GPU code */
static main_kernel_t staged_kernel(int threads, bool peer)
{
    if (peer) return threads <= 512 ? ehyb_staged_kernel<512, 4, true> : ehyb_staged_kernel<768, 4, true>;
    return threads <= 512 ? ehyb_staged_kernel<512, 4, false> : ehyb_staged_kernel<768, 4, false>;
}

/* persistent, double-buffered variant (ehyb_persistent_kernel): 4-column chunks only */
static main_kernel_t persistent_kernel(int threads, bool peer, int slots, bool dot = false)
{
    /* register budgets: 128 at <= 512 threads, 112 at <= 576, 96 at <= 640, 80 at 768.  The
     * multi-GPU build keeps a few more values live (calls into the exchange code): it spills at 80
     * registers, so peer sessions run at most kPeerPersistWarps = 20 warps.  Three staging slots
     * per warp exist for the 512-thread builds only (16 warps x 3 x 2.5 KB = 120 KB); the builds
     * with the fused dot product (ehyb_spmv_dot) have two slots.
     * $EHYB_PERSIST_BUILD (experiments) forces a build with a larger thread bound. */
    int b = env_int_early("EHYB_PERSIST_BUILD", 0);
    if (b < threads) b = threads;
    if (dot) {
        if (peer) return b <= 512 ? ehyb_persistent_kernel<512, 4, true, 2, true> : b <= 576 ? ehyb_persistent_kernel<576, 4, true, 2, true> : ehyb_persistent_kernel<640, 4, true, 2, true>;
        return b <= 512 ? ehyb_persistent_kernel<512, 4, false, 2, true> : b <= 576 ? ehyb_persistent_kernel<576, 4, false, 2, true>
               : b <= 640 ? ehyb_persistent_kernel<640, 4, false, 2, true> : ehyb_persistent_kernel<768, 4, false, 2, true>;
    }
    if (slots == 3) return peer ? ehyb_persistent_kernel<512, 4, true, 3> : ehyb_persistent_kernel<512, 4, false, 3>;
    if (peer) return b <= 512 ? ehyb_persistent_kernel<512, 4, true, 2> : b <= 576 ? ehyb_persistent_kernel<576, 4, true, 2> : ehyb_persistent_kernel<640, 4, true, 2>;
    return b <= 512 ? ehyb_persistent_kernel<512, 4, false, 2> : b <= 576 ? ehyb_persistent_kernel<576, 4, false, 2>
           : b <= 640 ? ehyb_persistent_kernel<640, 4, false, 2> : ehyb_persistent_kernel<768, 4, false, 2>;
}
constexpr int kPeerPersistWarps = 20;

/* The persistent kernel's work list.  seq = the partitions in `order` (NULL: identity).
 *   full rounds  CTA c takes whole partitions seq[c + G*j]: at any time the G CTAs stream G
 *                CONSECUTIVE partitions - one moving region of the blob and of x.  (Measured: giving
 *                every CTA its own contiguous run of the sequence instead, 148 streams spread over
 *                the whole blob, costs 7 % at 256^3 and 14 % at 512^3 / 2 GPUs.)
 *   tail         a whole number of partitions per CTA leaves SMs idle at the end (4 096 partitions
 *                on 148 SMs: 48 SMs idle for the last 1/28 of the product), and partitions differ in
 *                size.  The last partitions (the incomplete round; with >= 8 rounds also the last
 *                full one) are therefore cut by BYTES: CTA c gets a run of consecutive slices that
 *                tops its total up to the average, parts of at most a few partitions it shares
 *                with its neighbours.
 * Item = {rowStart, rowEnd, first slice, end slice (of the run inside the partition), cacheStart,
 * cacheCount, flags (bit 0: the cache list has halo columns), row of the item's first slice};
 * ctaStart[c] = first item of CTA c.  Multi-GPU sessions pass an order with the partitions that
 * have halo columns last: every CTA meets them at the end of its list. */
static int build_cta_tab(ehyb_handle *h, const ehyb_layout_view *v, const int32_t *order)
{
    const int P = h->nParts, G = h->grid;
    const int balance = env_int_early("EHYB_PERSIST_BALANCE", 1);
    int rounds = P / G;
    /* the tail: the P mod G partitions of the last, incomplete round - and the last full round as
     * well when a CTA has many rounds, to even out partitions of different sizes.  (A shared
     * partition is staged once per CTA that has a part of it: with 3 partitions per CTA - config 2,
     * 444 partitions - cutting a whole round of them cost 2.8 %, measured.) */
    if (balance && rounds >= 8) rounds -= 1;
    if (!balance) rounds = (P + G - 1) / G;           /* (experiments: whole partitions only) */
    const int tail0 = rounds * G < P ? rounds * G : P; /* first partition (sequence position) of the tail */
    const size_t maxItems = (size_t)P + 3 * (size_t)G + 1;
    int32_t *tab = (int32_t *)malloc(maxItems * 8 * sizeof(int32_t));
    int32_t *start = (int32_t *)malloc(((size_t)G + 1) * sizeof(int32_t));
    double *full = (double *)calloc((size_t)G, sizeof(double));
    if (!tab || !start || !full) { free(tab); free(start); free(full); return ehyb_fail(EHYB_ERR_NOMEM, "work list: out of memory"); }
    auto sliceCost = [&](int s2) { /* bytes streamed for the slice + its fixed cost (descriptor, y rows, chunk bookkeeping) */
        const ehyb_slice_desc &d = v->slices[s2];
        return (double)d.w * 512.0 + (double)((d.w + 3) / 4) * 512.0 + (double)d.wr * 512.0 + (double)((d.wr + 3) / 4) * 512.0 + 768.0;
    };
    auto partOf = [&](int i) { return &v->parts[order ? order[i] : i]; };
    auto emit = [&](int32_t *t, const ehyb_part_desc *d, int s0, int s1) {
        t[0] = d->rowStart; t[1] = d->rowEnd; t[2] = s0; t[3] = s1;
        t[4] = d->cacheStart; t[5] = d->cacheCount;
        t[6] = d->cacheCount > 0 && v->cacheCols[d->cacheStart + d->cacheCount - 1] >= v->n ? 1 : 0;
        t[7] = d->rowStart + (s0 - d->sliceStart) * EHYB_SLICE_ROWS;
    };
    double total = 0.0;
    for (int i = 0; i < P; ++i) {
        const ehyb_part_desc *d = partOf(i);
        double b = 0.0;
        for (int s2 = d->sliceStart; s2 < d->sliceEnd; ++s2) b += sliceCost(s2);
        total += b;
        if (i < tail0) full[i % G] += b;
    }
    const double avg = total / (double)G;
    size_t k = 0;
    int ti = tail0;                                        /* tail partition being cut */
    int sNext = ti < P ? partOf(ti)->sliceStart : 0;       /* its next unassigned slice */
    for (int c = 0; c < G; ++c) {
        start[c] = (int32_t)k;
        for (int j = 0; j < rounds && c + G * j < tail0; ++j) {
            const ehyb_part_desc *d = partOf(c + G * j);
            emit(tab + 8 * k, d, d->sliceStart, d->sliceEnd);
            ++k;
        }
        double have = full[c];
        while (ti < P) {
            const ehyb_part_desc *d = partOf(ti);
            if (sNext >= d->sliceEnd) { /* (also skips partitions without slices) */
                if (++ti < P) sNext = partOf(ti)->sliceStart;
                continue;
            }
            if (have >= avg && c + 1 < G) break;
            int sEnd = sNext;
            while (sEnd < d->sliceEnd && (c + 1 == G || have < avg)) have += sliceCost(sEnd++);
            emit(tab + 8 * k, d, sNext, sEnd);
            ++k;
            sNext = sEnd;
        }
    }
    start[G] = (int32_t)k;
    free(full);
    cudaError_t e = cudaSuccess;
    cudaFree(h->ctaTab); cudaFree(h->ctaStart);
    h->ctaTab = NULL; h->ctaStart = NULL;
    e = cudaMalloc(&h->ctaTab, (k ? k : 1) * 8 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&h->ctaStart, ((size_t)G + 1) * sizeof(int32_t));
    if (e == cudaSuccess && k) e = cudaMemcpy(h->ctaTab, tab, k * 8 * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->ctaStart, start, ((size_t)G + 1) * sizeof(int32_t), cudaMemcpyHostToDevice);
    free(tab); free(start);
    if (e != cudaSuccess) return ehyb_fail(EHYB_ERR_CUDA, "work list: %s", cudaGetErrorString(e));
    return EHYB_OK;
}

static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return s && s[0] ? atoi(s) : dflt;
}

extern "C" void ehyb_free(ehyb_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    cudaFree(h->parts); cudaFree(h->slices); cudaFree(h->blob);
    cudaFree(h->ovfRow); cudaFree(h->ovfCol); cudaFree(h->ovfVal);
    for (int b = 0; b < kMaxOvfBlocks; ++b) {
        cudaFree(h->ovfBlock[b].tiles); cudaFree(h->ovfBlock[b].rowOfSeg); cudaFree(h->ovfBlock[b].hubCols);
        cudaFree(h->ovfBlock[b].runs); cudaFree(h->ovfBlock[b].carryVal);
    }
    cudaFree(h->cacheCols); cudaFree(h->order); cudaFree(h->ctaTab); cudaFree(h->ctaStart); cudaFree(h->trace);
    cudaFree(h->x); cudaFree(h->y);
    for (int i = 0; i < 2; ++i) {
        cudaFree(h->xb[i]); cudaFree(h->yb[i]);
        if (h->evX[i]) cudaEventDestroy(h->evX[i]);
        if (h->evK[i]) cudaEventDestroy(h->evK[i]);
        if (h->evY[i]) cudaEventDestroy(h->evY[i]);
    }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->l2_persist) cudaCtxResetPersistingL2Cache();
    if (h->h2d) cudaStreamDestroy(h->h2d);
    if (h->d2h) cudaStreamDestroy(h->d2h);
    if (h->stream) cudaStreamDestroy(h->stream);
    free(h);
}

static int upload_impl(const ehyb_layout_view *v, const ehyb_session_opts *o, ehyb_handle *h, bool peerSession)
{
    CU(cudaSetDevice(o->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, o->device));
    if (prop.major < 9)
        return ehyb_fail(EHYB_ERR_CUDA, "device %d (%s, sm_%d%d) has no TMA bulk copies; this engine targets sm_100a", o->device,
                         prop.name, prop.major, prop.minor);
    h->device = o->device;
    h->smCount = prop.multiProcessorCount;
    h->n = v->n; h->ncols = v->ncols + (o->halo_cols > 0 && v->ncols == v->n ? o->halo_cols : 0);
    h->nnz = v->nnz; h->nOvf = v->nOverflow; h->blobBytes = v->blobBytes; h->algBytes = v->algBytes;
    h->nParts = v->nParts; h->W = v->W; h->kpp = v->ctasPerPart > 0 ? v->ctasPerPart : 1; h->nSlices = v->nSlices;
    /* default: the persistent kernel where the layout was planned for it (>= 3 partitions per SM,
     * one CTA per partition, room for >= 16 warps next to its two buffers), else the staged one */
    int kernel = o->kernel > 0 ? o->kernel : env_int("EHYB_KERNEL", 0);
    const int autoKernel = kernel <= 0;
    if (autoKernel) kernel = v->nParts >= 3 * prop.multiProcessorCount ? EHYB_KERNEL_PERSISTENT : EHYB_KERNEL_STAGED;
    int threads = o->threads > 0 ? o->threads : env_int("EHYB_THREADS", 0);
    const size_t winBytes = (((size_t)v->W + 2) * sizeof(double) + 127) & ~(size_t)127;
    h->cacheCap = (v->cacheMax + 15) & ~15;
    const size_t cacheBytes = ((size_t)h->cacheCap * sizeof(double) + 127) & ~(size_t)127;
    if (kernel == EHYB_KERNEL_PERSISTENT) {
        /* one CTA per SM over all of its partitions, {window, cache} double-buffered: needs one CTA
         * per partition and room for >= 8 warps of staging next to the two buffers (16 to be chosen
         * by default); otherwise the staged kernel */
        const int grid = h->nSlices < prop.multiProcessorCount ? (h->nSlices > 0 ? h->nSlices : 1) : prop.multiProcessorCount;
        const size_t fixed = (size_t)kPersistHeader + 2 * (winBytes + cacheBytes);
        /* Warps and staging slots (2.5 KB each), measured on 27-point grids (profiles/r2_notes.md):
         * what decides is the register budget of the consumer loop, not the bytes in flight -
         * 16 warps x 2 slots in the 128-register build (512 threads) stream 256^3 in 785 us, 20 x 2
         * at 96 registers in 828 us, 16 x 2 at 96 registers in 864 us, 16 x 3 slots in 802 us.  So:
         * 16 warps wherever more than 20 would fit; where the two buffers leave room for 17..20
         * (config 2: 20) all of them, which there is 1.5 % ahead of 16.  A third slot per warp is an
         * experiment switch ($EHYB_PERSIST_SLOTS=3, 512-thread builds only). */
        int slots = env_int("EHYB_PERSIST_SLOTS", 2);
        const size_t room = fixed < prop.sharedMemPerBlockOptin ? prop.sharedMemPerBlockOptin - fixed : 0;
        if (slots != 3 || room < (size_t)16 * 3 * slot_bytes(4) || (threads > 0 && threads != 512)) slots = 2;
        const size_t perWarp = (size_t)slots * slot_bytes(4);
        int nw = (int)(room / perWarp);
        if (nw > 20 && env_int("EHYB_PERSIST_WARPS", 0) <= 0) nw = 16;
        if (env_int("EHYB_PERSIST_WARPS", 0) > 0 && nw > env_int("EHYB_PERSIST_WARPS", 0)) nw = env_int("EHYB_PERSIST_WARPS", 0);
        if (nw > kMaxStageWarps) nw = kMaxStageWarps;
        if (slots == 3 && nw > 16) nw = 16;
        if ((peerSession || env_int("EHYB_FORCE_PEER_BUILD", 0) || env_int("EHYB_TRACE", 0)) && nw > kPeerPersistWarps) nw = kPeerPersistWarps;
        if (threads > 0 && threads / 32 < nw) nw = threads / 32 > 0 ? threads / 32 : 1;
        h->slotsPerWarp = slots;
        if (h->kpp != 1 || nw < (autoKernel ? 16 : 8)) {
            kernel = EHYB_KERNEL_STAGED;
        } else {
            h->kcEll = h->kcRem = 4;
            h->threads = nw * 32;
            h->smemBytes = fixed + (size_t)nw * perWarp;
            h->ctasPerSM = 1;
            h->grid = grid;
        }
    }
    if (kernel == EHYB_KERNEL_STAGED) {
        /* warps = staging capacity: 2 slots each, as many as fit next to the window */
        const int kc = 4;
        h->kcEll = kc;
        h->kcRem = kc;
        const size_t fixed = (size_t)kStageHeader + winBytes + cacheBytes;
        const size_t perWarp = (size_t)kSlotsPerWarp * slot_bytes(kc);
        int nw = fixed + perWarp <= prop.sharedMemPerBlockOptin ? (int)((prop.sharedMemPerBlockOptin - fixed) / perWarp) : 0;
        if (nw > kMaxStageWarps) nw = kMaxStageWarps;
        if (threads > 0 && threads / 32 < nw) nw = threads / 32 > 0 ? threads / 32 : 1;
        if (nw < 1) kernel = EHYB_KERNEL_DIRECT; /* window too large to stage next to it */
        else {
            h->threads = nw * 32;
            h->smemBytes = fixed + (size_t)nw * perWarp;
            h->ctasPerSM = (int)(prop.sharedMemPerMultiprocessor / (h->smemBytes + 1024));
            if (h->ctasPerSM < 1) h->ctasPerSM = 1;
        }
    }
    if (kernel != EHYB_KERNEL_STAGED && kernel != EHYB_KERNEL_PERSISTENT) {
        kernel = EHYB_KERNEL_DIRECT;
        h->smemBytes = (size_t)kSmemHeader + (size_t)((v->W + 2 + 15) & ~15) * sizeof(double) + cacheBytes;
        if (h->smemBytes > prop.sharedMemPerBlockOptin)
            return ehyb_fail(EHYB_ERR_LIMIT, "window of %d doubles needs %zu bytes of shared memory, device allows %zu", v->W,
                             h->smemBytes, (size_t)prop.sharedMemPerBlockOptin);
        /* resident CTAs per SM by shared memory (1 KB reserved per CTA) */
        int bySmem = (int)(prop.sharedMemPerMultiprocessor / (h->smemBytes + 1024));
        if (bySmem < 1) bySmem = 1;
        if (threads <= 0) threads = bySmem >= 2 ? 512 : 1024;
        threads = (threads + 31) / 32 * 32;
        if (threads > 1024) threads = 1024;
        h->threads = threads;
        h->ctasPerSM = bySmem;
        while (h->ctasPerSM > 1 && h->ctasPerSM * threads > 1024) h->ctasPerSM--;
    }
    h->kernel = kernel;
    if (kernel != EHYB_KERNEL_PERSISTENT) h->grid = h->nParts * h->kpp;

    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    CU(cudaMalloc(&h->parts, sizeof(ehyb_part_desc) * (size_t)h->nParts));
    CU(cudaMalloc(&h->slices, sizeof(ehyb_slice_desc) * (size_t)(h->nSlices ? h->nSlices : 1)));
    CU(cudaMalloc(&h->blob, (size_t)(h->blobBytes ? h->blobBytes : 256)));
    CU(cudaMemcpyAsync(h->parts, v->parts, sizeof(ehyb_part_desc) * (size_t)h->nParts, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->slices, v->slices, sizeof(ehyb_slice_desc) * (size_t)h->nSlices, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->blob, v->blob, (size_t)h->blobBytes, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMalloc(&h->cacheCols, sizeof(int32_t) * (size_t)(v->cacheTotal ? v->cacheTotal : 1)));
    if (v->cacheTotal) CU(cudaMemcpyAsync(h->cacheCols, v->cacheCols, sizeof(int32_t) * (size_t)v->cacheTotal, cudaMemcpyHostToDevice, h->stream));
    /* Overflow list.  Short lists (what the slices could not hold, halo entries) stay a COO reduced
     * with one atomic per row segment; long ones - >= $EHYB_OVF_STREAM_MIN entries, default 2^20, or
     * any length with $EHYB_DETERMINISTIC=1 - become the tile-packed CSR-like stream of
     * host/ovfstream.c: 12.4 instead of 16 bytes per entry, staged by bulk copies, hub columns in
     * shared memory, no atomics (bit-reproducible y).  Not for peer-memory sessions: there the
     * overflow kernel waits for the neighbours' flags.
     * Shape: 128-entry tiles x 32 warps (56 registers), four consecutive entries per lane: 4 096
     * gathers per SM in flight; $EHYB_OVF_SLOTS staging slots per warp (default 2) and $EHYB_OVF_HUBS
     * hub columns in shared memory (default 2 048).  Small shared-memory footprints on purpose: what is
     * not shared memory is L1, and on a power-law matrix the L1 hits of the warm columns are worth more
     * than a third slot or a large explicit hub cache (R-MAT 24, per product: 2 slots with 0 / 1 024 /
     * 2 048 / 4 096 / 6 144 / 8 192 / 12 288 hubs 1 599 / 1 498 / 1 469 / 1 535 / 1 511 / 1 851 / 2 964 us,
     * 3 slots without hubs 1 661; profiles/r2_notes.md). */
    h->ovfStream = 0;
    if (h->nOvf > 0 && !peerSession &&
        (env_int("EHYB_DETERMINISTIC", 0) || h->nOvf >= (int64_t)env_int("EHYB_OVF_STREAM_MIN", 1 << 20)) && env_int("EHYB_OVF_STREAM", 1)) {
        const int tg = 4;
        int slots = env_int("EHYB_OVF_SLOTS", 2);
        if (slots < 2) slots = 2;
        if (slots > kStreamMaxSlots) slots = kStreamMaxSlots;
        const int warps = tg == 8 ? 16 : 32;
        const size_t staging = (size_t)kStreamHeader + (size_t)warps * slots * EHYB_OVF_TILE_BYTES(tg);
        int hubCap = env_int("EHYB_OVF_HUBS", 2048);
        const int hubMax = staging < prop.sharedMemPerBlockOptin ? (int)((prop.sharedMemPerBlockOptin - staging) / sizeof(double)) : 0;
        if (hubCap > hubMax) hubCap = hubMax;
        if (hubCap < 0) hubCap = 0;
        /* column blocks ($EHYB_OVF_COLBLOCKS, default 1 = off): the list cut by column range and
         * streamed block after block, so that a launch's gathers fall into one slice of x.  Measured
         * on R-MAT 24 (x = 134 MB): 2 011 us with 1 block, 2 134 / 2 349 / 2 522 / 2 735 with 2 / 4 /
         * 8 / 16 - the gathers already hit L2 (86 %, ncu) and are bound by the L1TEX data pipe (one
         * wavefront per lane), so blocking only adds a pass over y per block.  Kept as a switch. */
        int nBlocks = env_int("EHYB_OVF_COLBLOCKS", 1);
        if (nBlocks > kMaxOvfBlocks) nBlocks = kMaxOvfBlocks;
        if (nBlocks < 1) nBlocks = 1;
        const int64_t blockCols = (h->ncols + nBlocks - 1) / nBlocks;
        ehyb_ovfstream st[kMaxOvfBlocks];
        int rcS = ehyb_ovfstream_build_blocked(h->nOvf, v->ovfRow, v->ovfCol, v->ovfVal, h->ncols, hubCap, tg, nBlocks, blockCols, st);
        if (rcS) return rcS;
        h->ovfStream = 1;
        h->ovfTileGroups = tg; h->ovfSlots = slots; h->ovfBlocks = nBlocks;
        h->ovfHubs = 0; h->ovfDeviceBytes = 0; h->ovfHubRefs = 0;
        cudaError_t e = cudaSuccess;
        for (int b = 0; b < nBlocks && e == cudaSuccess; ++b) {
            ehyb_handle::OvfBlock &ob = h->ovfBlock[b];
            ob.nTiles = (int)st[b].nTiles; ob.nHub = st[b].nHub;
            if (st[b].count == 0) continue;
            if (st[b].nHub > h->ovfHubs) h->ovfHubs = st[b].nHub;
            h->ovfDeviceBytes += st[b].deviceBytes; h->ovfHubRefs += st[b].hubRefs;
            const size_t blobBytes = (size_t)st[b].nTiles * (size_t)st[b].tileBytes;
            e = cudaMalloc(&ob.tiles, blobBytes);
            if (e == cudaSuccess) e = cudaMalloc(&ob.rowOfSeg, sizeof(int32_t) * (size_t)st[b].nSeg);
            if (e == cudaSuccess) e = cudaMalloc(&ob.hubCols, sizeof(int32_t) * (size_t)(st[b].nHub ? st[b].nHub : 1));
            ob.nRuns = st[b].nRuns; ob.nRunsShort = st[b].nRunsShort;
            if (e == cudaSuccess) e = cudaMalloc(&ob.runs, sizeof(int32_t) * 3 * (size_t)(st[b].nRuns ? st[b].nRuns : 1));
            if (e == cudaSuccess) e = cudaMalloc(&ob.carryVal, sizeof(double) * 2 * (size_t)ob.nTiles);
            if (e == cudaSuccess) e = cudaMemcpy(ob.tiles, st[b].tiles, blobBytes, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(ob.rowOfSeg, st[b].rowOfSeg, sizeof(int32_t) * (size_t)st[b].nSeg, cudaMemcpyHostToDevice);
            if (e == cudaSuccess && st[b].nHub) e = cudaMemcpy(ob.hubCols, st[b].hubCols, sizeof(int32_t) * (size_t)st[b].nHub, cudaMemcpyHostToDevice);
            /* the rows of the carry slots are fixed by the data (runs); the sums of unused slots stay 0 */
            if (e == cudaSuccess && st[b].nRuns) e = cudaMemcpy(ob.runs, st[b].runs, sizeof(int32_t) * 3 * (size_t)st[b].nRuns, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemset(ob.carryVal, 0, sizeof(double) * 2 * (size_t)ob.nTiles);
        }
        for (int b = 0; b < nBlocks; ++b) ehyb_ovfstream_free(&st[b]);
        if (e != cudaSuccess) return ehyb_fail(EHYB_ERR_CUDA, "overflow stream upload: %s", cudaGetErrorString(e));
        CU(cudaFuncSetAttribute(ehyb_ovfstream_kernel<4, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
    } else if (h->nOvf > 0) {
        CU(cudaMalloc(&h->ovfRow, sizeof(int32_t) * (size_t)h->nOvf));
        CU(cudaMalloc(&h->ovfCol, sizeof(int32_t) * (size_t)h->nOvf));
        CU(cudaMalloc(&h->ovfVal, sizeof(double) * (size_t)h->nOvf));
        CU(cudaMemcpyAsync(h->ovfRow, v->ovfRow, sizeof(int32_t) * (size_t)h->nOvf, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->ovfCol, v->ovfCol, sizeof(int32_t) * (size_t)h->nOvf, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->ovfVal, v->ovfVal, sizeof(double) * (size_t)h->nOvf, cudaMemcpyHostToDevice, h->stream));
    }
    CU(cudaMalloc(&h->x, sizeof(double) * (size_t)(h->ncols + 2)));
    CU(cudaMalloc(&h->y, sizeof(double) * (size_t)(h->n + 2)));
    CU(cudaMemsetAsync(h->x, 0, sizeof(double) * (size_t)(h->ncols + 2), h->stream));
    CU(cudaMemsetAsync(h->y, 0, sizeof(double) * (size_t)(h->n + 2), h->stream));

    /* shared-memory opt-in: once per session, not once per product (kernel.cu:351) */
    CU(cudaFuncSetAttribute(ehyb_main_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
    CU(cudaFuncSetAttribute(ehyb_main_kernel<1024, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
    CU(cudaFuncSetAttribute(ehyb_main_kernel<1024, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(ehyb_main_kernel<1024, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (int t = 512; t <= 768; t += 64)
        for (int peer = 0; peer < 2; ++peer)
            for (int sl = 2; sl <= 3; ++sl) {
                CU(cudaFuncSetAttribute(persistent_kernel(t, peer != 0, sl), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
                CU(cudaFuncSetAttribute(persistent_kernel(t, peer != 0, sl), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                CU(cudaFuncSetAttribute(persistent_kernel(t, peer != 0, sl, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
                CU(cudaFuncSetAttribute(persistent_kernel(t, peer != 0, sl, true), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            }
    if (kernel == EHYB_KERNEL_PERSISTENT) {
        int rcTab = build_cta_tab(h, v, NULL);
        if (rcTab) return rcTab;
    }
    for (int t = 512; t <= 768; t += 256)
        for (int peer = 0; peer < 2; ++peer) {
            CU(cudaFuncSetAttribute(staged_kernel(t, peer != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
            CU(cudaFuncSetAttribute(staged_kernel(t, peer != 0), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }

    h->use_graph = o->use_graph;
    h->pdl = env_int("EHYB_PDL", 1);
    h->forcePeerBuild = env_int("EHYB_FORCE_PEER_BUILD", 0);
    h->dbgSkip = env_int("EHYB_DEBUG_SKIP", 0); /* development: timing experiments without the arithmetic */
    h->haloInOverflow = v->haloInOverflow;
    /* nothing lives in slices (layout.c: coverage below min_coverage, everything in the COO list):
     * the product is a memset of y + the overflow kernel */
    h->skipMain = v->nnz > 0 && v->nnzEll + v->nnzRemInSlice == 0 && env_int("EHYB_SKIP_EMPTY_MAIN", 1);
    /* L2 eviction hints on the TMA copies: evict-first for the matrix stream and evict-last for x
     * when the matrix is larger than L2 (it would push x out every product: 100.8 -> 95.7 us at
     * config 2); none when matrix + vectors fit L2 and simply stay there (config 1: 60 MB) */
    h->l2hint = env_int("EHYB_L2_HINT", -1);
    if (h->l2hint < 0) h->l2hint = (double)v->blobBytes + 16.0 * (double)v->n > 0.75 * (double)prop.l2CacheSize;
    /* slices beyond the first round are dealt on demand only when a warp gets at least ~4 of them
     * (measured: neutral at config 2, +4 % on R-MAT with slices, -2 % at config 1 with 2.3) */
    h->dynamicDeal = env_int("EHYB_DYNAMIC_DEAL", -1);
    if (h->dynamicDeal < 0) h->dynamicDeal = (int64_t)h->nSlices >= 4 * (int64_t)h->nParts * h->kpp * (h->threads / 32);
    h->ovfUnroll = env_int("EHYB_OVF_UNROLL", 4);
    h->prologueBarrier = env_int("EHYB_PROLOGUE_BARRIER", 0);
    h->ovfLateTrigger = env_int("EHYB_OVF_LATE_TRIGGER", 1);
    h->pdlOvf = env_int("EHYB_PDL_OVF", 1); /* launch the overflow kernel programmatically behind the main kernel */
    h->winPiece = env_int("EHYB_WIN_PIECE", 32768) & ~15;
    if (h->winPiece < 16) h->winPiece = 32768;
    if (env_int("EHYB_TRACE", 0)) {
        const size_t traceCtas = (size_t)h->nParts * (size_t)h->kpp > (size_t)h->grid ? (size_t)h->nParts * (size_t)h->kpp : (size_t)h->grid;
        CU(cudaMalloc(&h->trace, sizeof(unsigned long long) * 8 * traceCtas));
        CU(cudaMemset(h->trace, 0, sizeof(unsigned long long) * 8 * traceCtas));
    }
    h->l2_persist = 0;
    if (o->l2_persist_x && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
        size_t want = sizeof(double) * (size_t)h->ncols;
        size_t carve = want < (size_t)prop.persistingL2CacheMaxSize ? want : (size_t)prop.persistingL2CacheMaxSize;
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof attr);
            size_t win = want < (size_t)prop.accessPolicyMaxWindowSize ? want : (size_t)prop.accessPolicyMaxWindowSize;
            attr.accessPolicyWindow.base_ptr = h->x;
            attr.accessPolicyWindow.num_bytes = win;
            attr.accessPolicyWindow.hitRatio = win <= carve ? 1.0f : (float)carve / (float)win;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess) h->l2_persist = 1;
        }
        cudaGetLastError();
    }
    CU(cudaStreamSynchronize(h->stream));
    return EHYB_OK;
}

static int upload_session(const ehyb_layout *L, const ehyb_session_opts *opts, bool peerSession, ehyb_handle **out);

extern "C" int ehyb_upload(const ehyb_layout *L, const ehyb_session_opts *opts, ehyb_handle **out)
{
    return upload_session(L, opts, false, out);
}

/* peerSession: the session will run the multi-GPU build of its kernel (peer-memory exchange) */
static int upload_session(const ehyb_layout *L, const ehyb_session_opts *opts, bool peerSession, ehyb_handle **out)
{
    if (!L || !out) return ehyb_fail(EHYB_ERR_ARG, "ehyb_upload: NULL argument");
    ehyb_session_opts o;
    if (opts) o = *opts;
    else ehyb_session_opts_default(&o);
    ehyb_layout_view v;
    int rc = ehyb_layout_get(L, &v);
    if (rc) return rc;
    ehyb_handle *h = (ehyb_handle *)calloc(1, sizeof *h);
    if (!h) return ehyb_fail(EHYB_ERR_NOMEM, "ehyb_upload: out of memory");
    rc = upload_impl(&v, &o, h, peerSession);
    if (rc) {
        char msg[512];
        snprintf(msg, sizeof msg, "%s", ehyb_last_error());
        ehyb_free(h);
        return ehyb_fail(rc, "%s", msg);
    }
    *out = h;
    return EHYB_OK;
}

/* peer: the launch carries a halo exchange, or the session records a per-CTA trace */
static main_kernel_t main_kernel_of(const ehyb_handle *h, bool peer, bool dot = false)
{
    if (h->kernel == EHYB_KERNEL_PERSISTENT) return persistent_kernel(h->threads, peer || h->trace != NULL || h->forcePeerBuild, h->slotsPerWarp, dot);
    if (h->kernel == EHYB_KERNEL_STAGED) return staged_kernel(h->threads, peer || h->trace != NULL || dot); /* the staged kernel's dot lives in its multi-GPU build */
    return pick_kernel(h->kernel, h->threads, h->ctasPerSM);
}

/* no exchange inside the kernels: halo columns (if any) are the tail of x */
static PeerArgs no_peer(const ehyb_handle *h, const double *x_d)
{
    PeerArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.xh = x_d + h->n;
    return pa;
}

static MainArgs main_args(const ehyb_handle *h, const double *x_d, double *y_d, const PeerArgs *pa)
{
    MainArgs a;
    a.dot = NULL;
    a.parts = h->parts; a.slices = h->slices; a.blob = h->blob;
    a.x = x_d; a.y = y_d; a.n = (int)h->n; a.W = h->W; a.kpp = h->kpp; a.dbg = h->dbgSkip; a.cacheCols = h->cacheCols; a.cacheCap = h->cacheCap;
    a.order = h->order;
    a.ctaTab = h->ctaTab;
    a.ctaStart = h->ctaStart;
    a.nPartsTotal = h->nParts;
    a.l2hint = h->l2hint;
    a.dynamicDeal = h->dynamicDeal;
    a.prologueBarrier = pa ? 0 : h->prologueBarrier;
    a.winPiece = (uint32_t)h->winPiece;
    a.trace = h->trace;
    a.peer = pa ? *pa : no_peer(h, x_d);
    return a;
}

/* the launches of one product, on `s` */
static int launch_main(ehyb_handle *h, const double *x_d, double *y_d, cudaStream_t s, const PeerArgs *pa, double *dot_d = NULL)
{
    if (h->skipMain && pa == NULL) { /* (a peer-memory product needs the kernel: it carries the halo push) */
        CU(cudaMemsetAsync(y_d, 0, sizeof(double) * (size_t)h->n, s));
        return EHYB_OK;
    }
    MainArgs a = main_args(h, x_d, y_d, pa);
    a.dot = dot_d;
    main_kernel_t k = main_kernel_of(h, pa != NULL, dot_d != NULL);
    if ((h->kernel == EHYB_KERNEL_STAGED || h->kernel == EHYB_KERNEL_PERSISTENT) && h->pdl) {
        /* programmatic dependent launch: this grid may start while the previous kernel of the
         * stream drains; it orders itself with griddepcontrol.wait before touching x or y */
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)h->grid);
        cfg.blockDim = dim3((unsigned)h->threads);
        cfg.dynamicSmemBytes = h->smemBytes;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        CU(cudaLaunchKernelEx(&cfg, k, a));
    } else {
        k<<<(unsigned)h->grid, h->threads, h->smemBytes, s>>>(a);
        CU(cudaGetLastError());
    }
    return EHYB_OK;
}

/* entries per warp of the overflow kernel: 32 while that still fits ~4 waves of the machine,
 * more (up to kOvfPerWarp) for long lists */
static int overflow_per_warp(const ehyb_handle *h)
{
    const int64_t warpsPerWave = (int64_t)h->smCount * 64;
    int64_t per = (h->nOvf + 4 * warpsPerWave - 1) / (4 * warpsPerWave);
    per = (per + 31) / 32 * 32;
    if (per < 32) per = 32;
    if (per > kOvfPerWarp) per = kOvfPerWarp;
    return (int)per;
}

static int launch_overflow(ehyb_handle *h, const double *x_d, double *y_d, cudaStream_t s, const PeerArgs *pa)
{
    if (h->nOvf <= 0) return EHYB_OK;
    if (h->ovfStream) {
        /* the CSR-like stream: one persistent CTA per SM, then the carry fix-up; plain launches (the
         * stream kernel reads x through the read-only path: it must not start before its predecessor
         * is complete) */
        if (pa != NULL && pa->flags != NULL) return ehyb_fail(EHYB_ERR_ARG, "the overflow stream does not carry the peer-memory exchange");
        const int tg = h->ovfTileGroups, warps = tg == 8 ? 16 : 32;
        for (int b = 0; b < h->ovfBlocks; ++b) {
            const ehyb_handle::OvfBlock &ob = h->ovfBlock[b];
            if (ob.nTiles == 0) continue;
            OvfStreamArgs a;
            a.tiles = ob.tiles; a.rowOfSeg = ob.rowOfSeg; a.hubCols = ob.hubCols;
            a.nHub = ob.nHub; a.nTiles = ob.nTiles; a.slots = h->ovfSlots;
            a.x = x_d; a.y = y_d; a.carryVal = ob.carryVal;
            /* y holds the slice part (main kernel) or zeros (memset); with several column blocks every
             * block adds to what the blocks before it left */
            a.accumulate = (h->skipMain && h->ovfBlocks == 1) ? 0 : 1;
            int grid = h->smCount;
            if ((int64_t)grid * warps > ob.nTiles) grid = (ob.nTiles + warps - 1) / warps; /* one tile per warp at least */
            const size_t smem = (size_t)kStreamHeader + (size_t)warps * h->ovfSlots * EHYB_OVF_TILE_BYTES(tg) + sizeof(double) * (size_t)ob.nHub;
            ehyb_ovfstream_kernel<4, 1024><<<grid, 1024, smem, s>>>(a);
            CU(cudaGetLastError());
            if (ob.nRuns > 0) {
                const int64_t blocks = (ob.nRunsShort + 255) / 256 + (ob.nRuns - ob.nRunsShort + 7) / 8;
                ehyb_ovfstream_fixup<<<(unsigned)blocks, 256, 0, s>>>(ob.runs, ob.nRunsShort, ob.nRuns, ob.carryVal, y_d, a.accumulate);
                CU(cudaGetLastError());
            }
        }
        return EHYB_OK;
    }
    OverflowArgs o;
    o.row = h->ovfRow; o.col = h->ovfCol; o.val = h->ovfVal; o.count = h->nOvf; o.x = x_d; o.y = y_d;
    o.perWarp = overflow_per_warp(h);
    o.n = (int)h->n;
    o.lateTrigger = h->ovfLateTrigger;
    o.peer = pa ? *pa : no_peer(h, x_d);
    const int64_t warps = (h->nOvf + o.perWarp - 1) / o.perWarp;
    const unsigned blocks = (unsigned)((warps + 7) / 8);
    /* Programmatic launch behind the main kernel pays for short lists (halo-sized: the launch gap
     * disappears), but with a grid of thousands of CTAs it costs 65 us per product (R-MAT scale
     * 20: 211 us vs 144, profiles/r1_notes.md): only for grids of at most 8 CTAs per SM. */
    if (h->pdl && h->pdlOvf && blocks <= 8u * (unsigned)h->smCount) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(blocks);
        cfg.blockDim = dim3(256);
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (h->ovfUnroll == 1) CU(cudaLaunchKernelEx(&cfg, ehyb_overflow_kernel<1>, o));
        else CU(cudaLaunchKernelEx(&cfg, ehyb_overflow_kernel<4>, o));
    } else {
        if (h->ovfUnroll == 1) ehyb_overflow_kernel<1><<<blocks, 256, 0, s>>>(o);
        else ehyb_overflow_kernel<4><<<blocks, 256, 0, s>>>(o);
        CU(cudaGetLastError());
    }
    return EHYB_OK;
}

static int launch_product(ehyb_handle *h, const double *x_d, double *y_d, cudaStream_t s)
{
    int rc = launch_main(h, x_d, y_d, s, NULL);
    return rc ? rc : launch_overflow(h, x_d, y_d, s, NULL);
}

extern "C" int ehyb_launches_per_spmv(const ehyb_handle *h)
{
    if (!h) return 0;
    int k = h->skipMain ? 0 : 1;
    if (h->nOvf > 0 && !h->ovfStream) k += 1;
    if (h->nOvf > 0 && h->ovfStream)
        for (int b = 0; b < h->ovfBlocks; ++b) /* stream kernel (+ carry fix-up) per column block */
            k += h->ovfBlock[b].nTiles > 0 ? (h->ovfBlock[b].nRuns > 0 ? 2 : 1) : 0;
    return k;
}

extern "C" int ehyb_spmv(ehyb_handle *h, const double *x_d, double *y_d)
{
    if (!h || !x_d || !y_d) return ehyb_fail(EHYB_ERR_ARG, "ehyb_spmv: NULL argument");
    CU(cudaSetDevice(h->device));
    if (!h->use_graph || h->nOvf == 0) return launch_product(h, x_d, y_d, h->stream); /* one launch: nothing to fuse */
    if (!h->gexec || h->gx != x_d || h->gy != y_d) {
        if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = NULL; }
        cudaGraph_t g;
        CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        int rc = launch_product(h, x_d, y_d, h->stream);
        cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        if (rc) return rc;
        if (e != cudaSuccess) return ehyb_fail(EHYB_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&h->gexec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return ehyb_fail(EHYB_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
        h->gx = x_d; h->gy = y_d;
    }
    CU(cudaGraphLaunch(h->gexec, h->stream));
    return EHYB_OK;
}

/* y = A x and *dot_d += x . y in one launch (the p.Ap of a conjugate-gradient iteration): the rows'
 * products are taken while y is stored, x[r] from the x window in shared memory.  Only where one
 * launch is the whole product: staged or persistent kernel, empty overflow list. */
extern "C" int ehyb_spmv_dot_supported(const ehyb_handle *h)
{
    return h && (h->kernel == EHYB_KERNEL_STAGED || h->kernel == EHYB_KERNEL_PERSISTENT) && h->nOvf == 0 && !h->skipMain &&
           (h->kernel == EHYB_KERNEL_STAGED || h->slotsPerWarp == 2);
}

extern "C" int ehyb_spmv_dot(ehyb_handle *h, const double *x_d, double *y_d, double *dot_d)
{
    if (!h || !x_d || !y_d || !dot_d) return ehyb_fail(EHYB_ERR_ARG, "ehyb_spmv_dot: NULL argument");
    if (!ehyb_spmv_dot_supported(h)) return ehyb_fail(EHYB_ERR_ARG, "ehyb_spmv_dot: this session's product is not a single staged / persistent launch");
    CU(cudaSetDevice(h->device));
    return launch_main(h, x_d, y_d, h->stream, NULL, dot_d);
}

/* host-vector form (tests, one-off callers): y_h = A x_h, *dot_h = x_h . y_h from the fused kernel */
extern "C" int ehyb_spmv_dot_host(ehyb_handle *h, const double *x_h, double *y_h, double *dot_h)
{
    if (!h || !x_h || !y_h || !dot_h) return ehyb_fail(EHYB_ERR_ARG, "ehyb_spmv_dot_host: NULL argument");
    CU(cudaSetDevice(h->device));
    double *dot_d = NULL;
    CU(cudaMalloc(&dot_d, sizeof(double)));
    int rc = EHYB_OK;
    cudaError_t e = cudaMemsetAsync(dot_d, 0, sizeof(double), h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->x, x_h, sizeof(double) * (size_t)h->n, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) rc = ehyb_spmv_dot(h, h->x, h->y, dot_d);
    if (e == cudaSuccess && rc == EHYB_OK) e = cudaMemcpyAsync(y_h, h->y, sizeof(double) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && rc == EHYB_OK) e = cudaMemcpyAsync(dot_h, dot_d, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dot_d);
    if (rc) return rc;
    if (e != cudaSuccess) return ehyb_fail(EHYB_ERR_CUDA, "ehyb_spmv_dot_host: %s", cudaGetErrorString(e));
    return EHYB_OK;
}

extern "C" int ehyb_sync(ehyb_handle *h)
{
    if (!h) return ehyb_fail(EHYB_ERR_ARG, "ehyb_sync: NULL");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return peer_check(h);
}

extern "C" void *ehyb_stream(ehyb_handle *h) { return h ? (void *)h->stream : NULL; }

extern "C" int ehyb_session_vectors(ehyb_handle *h, double **x_d, double **y_d)
{
    if (!h) return ehyb_fail(EHYB_ERR_ARG, "ehyb_session_vectors: NULL");
    if (x_d) *x_d = h->x;
    if (y_d) *y_d = h->y;
    return EHYB_OK;
}

extern "C" int ehyb_set_x(ehyb_handle *h, const double *x_h)
{
    if (!h || !x_h) return ehyb_fail(EHYB_ERR_ARG, "ehyb_set_x: NULL");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->x, x_h, sizeof(double) * (size_t)h->ncols, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return EHYB_OK;
}

extern "C" int ehyb_get_y(ehyb_handle *h, double *y_h)
{
    if (!h || !y_h) return ehyb_fail(EHYB_ERR_ARG, "ehyb_get_y: NULL");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(y_h, h->y, sizeof(double) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return peer_check(h);
}

extern "C" int ehyb_spmv_host(ehyb_handle *h, const double *x_h, double *y_h)
{
    if (!h || !x_h || !y_h) return ehyb_fail(EHYB_ERR_ARG, "ehyb_spmv_host: NULL argument");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->x, x_h, sizeof(double) * (size_t)h->ncols, cudaMemcpyHostToDevice, h->stream));
    int rc = ehyb_spmv(h, h->x, h->y);
    if (rc) return rc;
    CU(cudaMemcpyAsync(y_h, h->y, sizeof(double) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return EHYB_OK;
}

static int ensure_pipeline(ehyb_handle *h)
{
    if (h->h2d) return EHYB_OK;
    CU(cudaStreamCreateWithFlags(&h->h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CU(cudaMalloc(&h->xb[i], sizeof(double) * (size_t)(h->ncols + 2)));
        CU(cudaMalloc(&h->yb[i], sizeof(double) * (size_t)(h->n + 2)));
        CU(cudaEventCreateWithFlags(&h->evX[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->evK[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->evY[i], cudaEventDisableTiming));
    }
    return EHYB_OK;
}

extern "C" int ehyb_spmv_host_batch(ehyb_handle *h, const double *const *x_h, double *const *y_h, int count)
{
    if (!h || !x_h || !y_h || count < 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_spmv_host_batch: bad argument");
    CU(cudaSetDevice(h->device));
    int rc = ensure_pipeline(h);
    if (rc) return rc;
    /* product i uses buffer pair i%2: its H2D waits for kernel i-2, its kernel for D2H i-2 */
    for (int i = 0; i < count; ++i) {
        const int b = i & 1;
        if (i >= 2) CU(cudaStreamWaitEvent(h->h2d, h->evK[b], 0));
        CU(cudaMemcpyAsync(h->xb[b], x_h[i], sizeof(double) * (size_t)h->ncols, cudaMemcpyHostToDevice, h->h2d));
        CU(cudaEventRecord(h->evX[b], h->h2d));
        CU(cudaStreamWaitEvent(h->stream, h->evX[b], 0));
        if (i >= 2) CU(cudaStreamWaitEvent(h->stream, h->evY[b], 0));
        rc = launch_product(h, h->xb[b], h->yb[b], h->stream);
        if (rc) return rc;
        CU(cudaEventRecord(h->evK[b], h->stream));
        CU(cudaStreamWaitEvent(h->d2h, h->evK[b], 0));
        CU(cudaMemcpyAsync(y_h[i], h->yb[b], sizeof(double) * (size_t)h->n, cudaMemcpyDeviceToHost, h->d2h));
        CU(cudaEventRecord(h->evY[b], h->d2h));
    }
    CU(cudaStreamSynchronize(h->h2d));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaStreamSynchronize(h->d2h));
    return EHYB_OK;
}

extern "C" int ehyb_time_spmv(ehyb_handle *h, int warmup, int iters, float *ms_total, float *kernel_ms)
{
    if (!h || !ms_total || iters <= 0 || warmup < 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_time_spmv: bad argument");
    CU(cudaSetDevice(h->device));
    for (int i = 0; i < warmup; ++i) {
        int rc = ehyb_spmv(h, h->x, h->y);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaEventRecord(h->ev0, h->stream));
    for (int i = 0; i < iters; ++i) {
        int rc = ehyb_spmv(h, h->x, h->y);
        if (rc) return rc;
    }
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaEventElapsedTime(ms_total, h->ev0, h->ev1));
    if (kernel_ms) {
        /* main kernel alone: one event pair per launch, summed */
        const MainArgs a = main_args(h, h->x, h->y, NULL);
        main_kernel_t k = main_kernel_of(h, false);
        cudaEvent_t *ev = (cudaEvent_t *)calloc((size_t)iters * 2, sizeof(cudaEvent_t));
        if (!ev) return ehyb_fail(EHYB_ERR_NOMEM, "ehyb_time_spmv: out of memory");
        for (int i = 0; i < 2 * iters; ++i) cudaEventCreate(&ev[i]);
        for (int i = 0; i < iters; ++i) {
            cudaEventRecord(ev[2 * i], h->stream);
            if (h->skipMain) cudaMemsetAsync(h->y, 0, sizeof(double) * (size_t)h->n, h->stream);
            else k<<<(unsigned)h->grid, h->threads, h->smemBytes, h->stream>>>(a);
            cudaEventRecord(ev[2 * i + 1], h->stream);
            launch_overflow(h, h->x, h->y, h->stream, NULL);
        }
        cudaError_t e = cudaStreamSynchronize(h->stream);
        float sum = 0.f;
        for (int i = 0; i < iters && e == cudaSuccess; ++i) {
            float t = 0.f;
            e = cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]);
            sum += t;
        }
        for (int i = 0; i < 2 * iters; ++i) cudaEventDestroy(ev[i]);
        free(ev);
        if (e != cudaSuccess) return ehyb_fail(EHYB_ERR_CUDA, "kernel timing: %s", cudaGetErrorString(e));
        *kernel_ms = sum;
    }
    return EHYB_OK;
}

/* Cold-L2 timing (SURVEY.md 8d "Timing": matrices that fit the 126 MB L2 must be reported with and
 * without a flush): before every product `flush_bytes` of scratch memory are overwritten on the
 * session stream, which pushes matrix, x and y out of L2; every product has its own event pair (all
 * of its launches between them) and *ms_sum is the sum over `iters` products.  Launches are plain
 * here: a programmatic launch would start the product's prologue under the flush. */
extern "C" int ehyb_time_spmv_flushed(ehyb_handle *h, int warmup, int iters, size_t flush_bytes, float *ms_sum)
{
    if (!h || !ms_sum || iters <= 0 || warmup < 0 || flush_bytes == 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_time_spmv_flushed: bad argument");
    CU(cudaSetDevice(h->device));
    void *scratch = NULL;
    CU(cudaMalloc(&scratch, flush_bytes));
    cudaEvent_t *ev = (cudaEvent_t *)calloc((size_t)iters * 2, sizeof(cudaEvent_t));
    if (!ev) { cudaFree(scratch); return ehyb_fail(EHYB_ERR_NOMEM, "ehyb_time_spmv_flushed: out of memory"); }
    for (int i = 0; i < 2 * iters; ++i) cudaEventCreate(&ev[i]);
    const int pdl = h->pdl;
    h->pdl = 0;
    int rc = EHYB_OK;
    for (int i = -warmup; i < iters && rc == EHYB_OK; ++i) {
        cudaMemsetAsync(scratch, i & 1, flush_bytes, h->stream);
        if (i >= 0) cudaEventRecord(ev[2 * i], h->stream);
        rc = launch_product(h, h->x, h->y, h->stream);
        if (i >= 0) cudaEventRecord(ev[2 * i + 1], h->stream);
    }
    h->pdl = pdl;
    cudaError_t e = cudaStreamSynchronize(h->stream);
    float sum = 0.f;
    for (int i = 0; i < iters && e == cudaSuccess && rc == EHYB_OK; ++i) {
        float t = 0.f;
        e = cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]);
        sum += t;
    }
    for (int i = 0; i < 2 * iters; ++i) cudaEventDestroy(ev[i]);
    free(ev);
    cudaFree(scratch);
    if (rc) return rc;
    if (e != cudaSuccess) return ehyb_fail(EHYB_ERR_CUDA, "flushed timing: %s", cudaGetErrorString(e));
    *ms_sum = sum;
    return EHYB_OK;
}

/* Development aid (EHYB_TRACE=1 when the session was created): the staged kernel's per-CTA
 * timeline of the last product - 8 words per CTA: globaltimer ns at CTA start, after the wait
 * for the previous grid, after the halo push, when window + cache are staged, when the last
 * warp finished; then SM id, partition, unused.  out holds 8 * *ctas words. */
extern "C" int ehyb_trace_read(ehyb_handle *h, unsigned long long *out, int *ctas)
{
    if (!h || !ctas) return ehyb_fail(EHYB_ERR_ARG, "ehyb_trace_read: NULL");
    *ctas = h->trace ? (h->nParts * h->kpp > h->grid ? h->nParts * h->kpp : h->grid) : 0;
    if (!h->trace || !out) return EHYB_OK;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(out, h->trace, sizeof(unsigned long long) * 8 * (size_t)*ctas, cudaMemcpyDeviceToHost));
    return EHYB_OK;
}

extern "C" const char *ehyb_session_kernel(const ehyb_handle *h)
{
    if (!h) return "";
    if (h->skipMain) return h->ovfStream ? "ehyb_ovfstream_kernel" : "ehyb_overflow_kernel";
    return h->kernel == EHYB_KERNEL_PERSISTENT ? "ehyb_persistent_kernel" : h->kernel == EHYB_KERNEL_STAGED ? "ehyb_staged_kernel" : "ehyb_main_kernel";
}

extern "C" int ehyb_session_size(const ehyb_handle *h, int64_t *n, int64_t *ncols)
{
    if (!h) return ehyb_fail(EHYB_ERR_ARG, "ehyb_session_size: NULL");
    if (n) *n = h->n;
    if (ncols) *ncols = h->ncols;
    return EHYB_OK;
}

extern "C" int ehyb_describe(ehyb_handle *h, matrixEHYB *d)
{
    if (!h || !d) return ehyb_fail(EHYB_ERR_ARG, "ehyb_describe: NULL");
    memset(d, 0, sizeof *d);
    d->dimension = (int)h->n;
    d->nParts = h->nParts;
    d->vectorCacheSize = (int16_t)h->W;
    d->kernelPerPart = h->kpp;
    d->b200 = h;
    return EHYB_OK;
}

extern "C" int ehyb_session_info(const ehyb_handle *h, int *threads, int *ctasPerSM, int *grid, int64_t *smemBytes, int *l2_persist)
{
    if (!h) return ehyb_fail(EHYB_ERR_ARG, "ehyb_session_info: NULL");
    if (threads) *threads = h->threads;
    if (ctasPerSM) *ctasPerSM = h->ctasPerSM;
    if (grid) *grid = h->grid;
    if (smemBytes) *smemBytes = (int64_t)h->smemBytes;
    if (l2_persist) *l2_persist = h->l2_persist;
    return EHYB_OK;
}

/* pinned host memory for callers that want true asynchronous copies */
extern "C" int ehyb_host_alloc_pinned(size_t bytes, void **out)
{
    if (!out) return ehyb_fail(EHYB_ERR_ARG, "ehyb_host_alloc_pinned: NULL");
    CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return EHYB_OK;
}

extern "C" int ehyb_host_free_pinned(void *p)
{
    CU(cudaFreeHost(p));
    return EHYB_OK;
}

/* ====================================================================================== */
/* multi-GPU: one process per GPU, x halo exchanged with NCCL send/recv every product       */
/* ====================================================================================== */
#include <dlfcn.h>
#include <nccl.h>

/* host side (host/mg.c) */
extern "C" int ehyb_mg_local_view(const ehyb_mg_local *L, const matrixCOO **coo, const ehyb_layout **layout, int64_t *nSend,
                                  const int32_t **sendIdx, const int64_t **sendCount);
extern "C" int ehyb_mg_local_halo(const ehyb_mg_local *L, int64_t *nHalo, const int64_t **haloGlobal, const int64_t **recvCount);

/* NCCL is bound at run time (the library must load on hosts without it; under PyTorch the
 * process-wide libnccl.so.2 is the one torch already loaded) */
struct NcclApi {
    void *dl;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

static int nccl_load(void)
{
    if (g_nccl.dl) return EHYB_OK;
    const char *names[] = {getenv("EHYB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *dl = NULL;
    for (int i = 0; i < 3 && !dl; ++i)
        if (names[i] && names[i][0]) dl = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!dl) return ehyb_fail(EHYB_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                  \
    *(void **)(&g_nccl.field) = dlsym(dl, name);                                          \
    if (!g_nccl.field) return ehyb_fail(EHYB_ERR_NCCL, "libnccl lacks %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.dl = dl;
    return EHYB_OK;
}

#define NC(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess)                                                                           \
            return ehyb_fail(EHYB_ERR_NCCL, "%s: %s (%s:%d)", #call, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

struct ehyb_mg_session {
    ehyb_handle *h;
    int rank, nranks, exchange;
    int64_t nSend, nHalo;
    int32_t *sendIdx_d;
    int64_t *sendCount, *recvCount; /* host copies */
    /* NCCL exchange */
    ncclComm_t comm;
    cudaStream_t commStream;
    cudaEvent_t evX, evHalo;
    double *sendBuf_d;
    /* peer-memory exchange */
    unsigned char *shared;       /* one allocation mapped by the neighbours: [halo 0 | halo 1 | flags] */
    size_t sharedBytes, haloStride; /* haloStride: bytes between the two halo buffers */
    void **peerBase;             /* [nranks] the neighbours' `shared`, mapped here (NULL: not a neighbour) */
    double **pushDst_d[2];       /* device: [nSend] destination of every send-list entry, per parity */
    uint32_t **peerFlag_d;       /* device: [nPeers] flags[my rank][0] on every neighbour */
    int32_t *peerPushCtas_d;     /* device: [nranks] pushing CTAs of every rank (flag words to poll) */
    uint32_t *status_d;          /* device view of status_h */
    uint32_t *status_h;          /* mapped pinned host word, see ehyb_handle.peerStatus_h */
    uint32_t epoch, recvMask, nbrMask;
    int nPeers, pushCtas, connected, peerIsLocal;
    unsigned long long timeoutNs;
    /* scalar all-reduce over peer memory (ehyb_mg_allreduce_sum): a mailbox in `shared` that EVERY
     * rank maps - [2 parities][nranks][kMboxVals] doubles, then [2][nranks] epoch words */
    size_t mboxOffset;
    unsigned char **mboxPeer_d;  /* device: [nranks] every rank's mailbox (own: local address) */
    uint32_t arEpoch;
};
constexpr int kMboxVals = 4;
static size_t mbox_bytes(int nranks) { return ((size_t)2 * nranks * kMboxVals * sizeof(double) + (size_t)2 * nranks * sizeof(uint32_t) + 255) & ~(size_t)255; }

__global__ void ehyb_pack_kernel(const double *__restrict__ x, const int32_t *__restrict__ idx, double *__restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = x[idx[i]];
}

extern "C" int ehyb_mg_unique_id(void *id128)
{
    if (!id128) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_unique_id: NULL");
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return EHYB_OK;
}

extern "C" void ehyb_mg_session_free(ehyb_mg_session *s)
{
    if (!s) return;
    if (s->h) {
        cudaSetDevice(s->h->device);
        if (s->h->stream) cudaStreamSynchronize(s->h->stream);
    }
    if (s->commStream) cudaStreamSynchronize(s->commStream);
    if (s->comm && g_nccl.dl) g_nccl.CommDestroy(s->comm);
    if (s->peerBase) {
        for (int g = 0; g < s->nranks; ++g)
            if (s->peerBase[g] && !s->peerIsLocal) cudaIpcCloseMemHandle(s->peerBase[g]);
        free(s->peerBase);
    }
    cudaFree(s->shared); cudaFree(s->pushDst_d[0]); cudaFree(s->pushDst_d[1]); cudaFree(s->peerFlag_d); cudaFree(s->mboxPeer_d);
    cudaFree(s->peerPushCtas_d);
    if (s->status_h) { if (s->h) s->h->peerStatus_h = NULL; cudaFreeHost(s->status_h); }
    cudaFree(s->sendIdx_d); cudaFree(s->sendBuf_d);
    if (s->evX) cudaEventDestroy(s->evX);
    if (s->evHalo) cudaEventDestroy(s->evHalo);
    if (s->commStream) cudaStreamDestroy(s->commStream);
    free(s->sendCount); free(s->recvCount);
    ehyb_free(s->h);
    free(s);
}

/* what both exchanges share: the uploaded block, the send list and the per-peer counts */
static int mg_session_base(const ehyb_mg_local *L, int rank, int nranks, int device, int exchange, ehyb_mg_session **out)
{
    const ehyb_layout *layout = NULL;
    const int32_t *sendIdx = NULL;
    const int64_t *sendCount = NULL, *recvCount = NULL;
    int64_t nSend = 0, nHalo = 0;
    int rc = ehyb_mg_local_view(L, NULL, &layout, &nSend, &sendIdx, &sendCount);
    if (rc) return rc;
    ehyb_mg_local_halo(L, &nHalo, NULL, &recvCount);
    ehyb_layout_view v;
    rc = ehyb_layout_get(layout, &v);
    if (rc) return rc;
    if (nHalo > 0 && (v.haloInOverflow != 0) != (exchange == EHYB_MG_NCCL))
        return ehyb_fail(EHYB_ERR_ARG, "the block was finished for the %s exchange", v.haloInOverflow ? "NCCL" : "peer-memory");
    ehyb_mg_session *s = (ehyb_mg_session *)calloc(1, sizeof *s);
    if (!s) return ehyb_fail(EHYB_ERR_NOMEM, "mg session: out of memory");
    s->rank = rank; s->nranks = nranks; s->nSend = nSend; s->nHalo = nHalo; s->exchange = exchange;
    ehyb_session_opts o;
    ehyb_session_opts_default(&o);
    o.device = device;
    /* the halo push and pull live in the persistent and in the staged kernel; the direct one is
     * not a choice here.  Default (0): persistent where the layout allows it ($EHYB_MG_KERNEL) */
    o.kernel = env_int("EHYB_MG_KERNEL", 0);
    if (o.kernel == EHYB_KERNEL_DIRECT) o.kernel = EHYB_KERNEL_STAGED;
    rc = upload_session(layout, &o, exchange == EHYB_MG_P2P, &s->h);
    if (rc) { free(s); return rc; }
    auto body = [&]() -> int {
        CU(cudaSetDevice(device));
        s->sendCount = (int64_t *)malloc(sizeof(int64_t) * (size_t)nranks);
        s->recvCount = (int64_t *)malloc(sizeof(int64_t) * (size_t)nranks);
        if (!s->sendCount || !s->recvCount) return ehyb_fail(EHYB_ERR_NOMEM, "mg session: out of memory");
        memcpy(s->sendCount, sendCount, sizeof(int64_t) * (size_t)nranks);
        memcpy(s->recvCount, recvCount, sizeof(int64_t) * (size_t)nranks);
        CU(cudaMalloc(&s->sendIdx_d, sizeof(int32_t) * (size_t)(nSend ? nSend : 1)));
        if (nSend) CU(cudaMemcpy(s->sendIdx_d, sendIdx, sizeof(int32_t) * (size_t)nSend, cudaMemcpyHostToDevice));
        return EHYB_OK;
    };
    rc = body();
    if (rc) {
        char msg[512];
        snprintf(msg, sizeof msg, "%s", ehyb_last_error());
        ehyb_mg_session_free(s);
        return ehyb_fail(rc, "%s", msg);
    }
    *out = s;
    return EHYB_OK;
}

#define MG_TRY(s, call)                                            \
    do {                                                           \
        int r2__ = (call);                                         \
        if (r2__) {                                                \
            char m__[512];                                         \
            snprintf(m__, sizeof m__, "%s", ehyb_last_error());    \
            ehyb_mg_session_free(s);                               \
            return ehyb_fail(r2__, "%s", m__);                     \
        }                                                          \
    } while (0)

/* Collective over all ranks (ncclCommInitRank).  id128 comes from rank 0's ehyb_mg_unique_id,
 * distributed by the caller. */
extern "C" int ehyb_mg_session_create(const ehyb_mg_local *L, int rank, int nranks, int device, const void *id128,
                                      ehyb_mg_session **out)
{
    if (!L || !id128 || !out || nranks <= 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_session_create: bad argument");
    int rc = nccl_load();
    if (rc) return rc;
    ehyb_mg_session *s = NULL;
    rc = mg_session_base(L, rank, nranks, device, EHYB_MG_NCCL, &s);
    if (rc) return rc;
    auto body = [&]() -> int {
        CU(cudaStreamCreateWithFlags(&s->commStream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s->evX, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->evHalo, cudaEventDisableTiming));
        CU(cudaMalloc(&s->sendBuf_d, sizeof(double) * (size_t)(s->nSend ? s->nSend : 1)));
        ncclUniqueId id;
        memcpy(&id, id128, sizeof id);
        NC(g_nccl.CommInitRank(&s->comm, nranks, id, rank));
        return EHYB_OK;
    };
    MG_TRY(s, body());
    *out = s;
    return EHYB_OK;
}

/* ---- peer-memory exchange ---------------------------------------------------------------- */

/* what a rank tells the others about its shared allocation (EHYB_MG_P2P_BLOB_BYTES) */
struct P2PBlob {
    cudaIpcMemHandle_t handle; /* 64 bytes */
    int64_t haloStride;        /* bytes from halo buffer 0 to halo buffer 1 */
    int64_t flagsOffset;       /* bytes from the base to flags[0] */
    int64_t nHalo;
    int32_t device, rank;
    int32_t pushCtas;          /* CTAs of this rank that push = flag words it writes on a neighbour */
    int32_t pad32;
    int64_t mboxOffset;        /* bytes from the base to the all-reduce mailbox */
    int64_t pad[2];
};
static_assert(sizeof(P2PBlob) == EHYB_MG_P2P_BLOB_BYTES, "P2PBlob must match EHYB_MG_P2P_BLOB_BYTES");

extern "C" int ehyb_mg_p2p_supported(int device, int nranks, int *supported)
{
    if (!supported || nranks <= 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_p2p_supported: bad argument");
    *supported = 0;
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (nranks > 32) return EHYB_OK; /* flags masks are 32-bit */
    if (nranks > count) {
        /* more ranks than GPUs: only for tests ($EHYB_MG_SHARE_DEVICE=1) - several ranks (processes) share a
         * GPU, reach each other's buffers through CUDA IPC on the SAME device, and their kernels take
         * turns on it (time slicing): the exchange protocol runs on a one-GPU box, at no useful speed */
        *supported = env_int("EHYB_MG_SHARE_DEVICE", 0) != 0 && count >= 1;
        return EHYB_OK;
    }
    for (int g = 0; g < nranks; ++g) {
        if (g == device) continue;
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, device, g));
        if (!can) return EHYB_OK;
    }
    *supported = 1;
    return EHYB_OK;
}

extern "C" int ehyb_mg_session_create_p2p(const ehyb_mg_local *L, int rank, int nranks, int device, ehyb_mg_session **out)
{
    if (!L || !out || nranks <= 0 || nranks > 32) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_session_create_p2p: bad argument (at most 32 ranks)");
    ehyb_mg_session *s = NULL;
    int rc = mg_session_base(L, rank, nranks, device, EHYB_MG_P2P, &s);
    if (rc) return rc;
    const ehyb_layout *layout = NULL;
    ehyb_mg_local_view(L, NULL, &layout, NULL, NULL, NULL);
    auto body = [&]() -> int {
        /* [halo 0 | halo 1 | flags], every part 256-byte aligned; the allocation is rounded up to
         * 2 MiB so that the IPC handle covers this block and nothing else */
        s->haloStride = (((size_t)s->nHalo + 1) * sizeof(double) + 255) & ~(size_t)255;
        const size_t flagsBytes = ((size_t)nranks * kMaxPushCtas * sizeof(uint32_t) + 255) & ~(size_t)255;
        s->mboxOffset = 2 * s->haloStride + flagsBytes;
        s->sharedBytes = (s->mboxOffset + mbox_bytes(nranks) + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
        CU(cudaMalloc(&s->shared, s->sharedBytes));
        CU(cudaMemset(s->shared, 0, s->sharedBytes));
        CU(cudaHostAlloc((void **)&s->status_h, 256, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(s->status_h, 0, 256);
        CU(cudaHostGetDevicePointer((void **)&s->status_d, s->status_h, 0));
        s->h->peerStatus_h = s->status_h;
        CU(cudaDeviceSynchronize());
        /* time limit of every wait on a neighbour = bound on the launch skew between the ranks
         * (0: no limit) */
        const int toMs = env_int("EHYB_P2P_TIMEOUT_MS", 10000);
        s->timeoutNs = toMs > 0 ? (unsigned long long)toMs * 1000000ull : 0ull;
        /* the CTAs that push (their last warp does) are resident in the first wave of the main
         * kernel: at most one per SM and no more than the grid; >= 64 entries each */
        ehyb_handle *h = s->h;
        if (h->kernel == EHYB_KERNEL_DIRECT) return ehyb_fail(EHYB_ERR_LIMIT, "the window leaves no room for the staging slots: no kernel with the halo exchange fits");
        int64_t ctas = (s->nSend + 63) / 64;
        if (env_int("EHYB_P2P_PUSH_CTAS", 0) > 0) ctas = env_int("EHYB_P2P_PUSH_CTAS", 0);
        if (ctas > h->grid) ctas = h->grid;
        if (ctas > h->smCount) ctas = h->smCount;
        if (ctas > kMaxPushCtas) ctas = kMaxPushCtas;
        if (ctas < 1) ctas = 1;
        s->pushCtas = (int)ctas;
        /* Dispatch order.  Partitions WITH halo columns in their remainder cache (the lists are
         * ascending, so the last entry tells) wait for the neighbours' push of this product; the
         * CTAs that push must not.
         *   staged kernel (CTAs start in blockIdx order): a first wave of partitions without halo
         *     columns - these CTAs push, nobody in them waits -, then the partitions that need halo
         *     values, by which time the neighbours' push has arrived, then the rest, so that the
         *     tail of the kernel is made of ordinary partitions;
         *   persistent kernel (CTA c takes the partitions order[c + grid*j], build_cta_tab): all the
         *     partitions without halo columns first, the others last - every CTA meets them at the
         *     end of its list, when the neighbours have long delivered. */
        ehyb_layout_view v;
        int rc2 = ehyb_layout_get(layout, &v);
        if (rc2) return rc2;
        int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)h->nParts);
        unsigned char *cls = (unsigned char *)malloc((size_t)h->nParts);
        if (!order || !cls) { free(order); free(cls); return ehyb_fail(EHYB_ERR_NOMEM, "mg session: out of memory"); }
        const int reorder = env_int("EHYB_P2P_ORDER", 1);
        int firstWave = h->smCount * h->ctasPerSM / h->kpp;
        if (firstWave < 1) firstWave = 1;
        int nPlain = 0;
        for (int p = 0; p < h->nParts; ++p) {
            const int cnt = v.parts[p].cacheCount;
            const int halo = reorder && cnt > 0 && v.cacheCols[v.parts[p].cacheStart + cnt - 1] >= v.n;
            cls[p] = halo ? 1 : (h->kernel == EHYB_KERNEL_PERSISTENT || nPlain++ < firstWave ? 0 : 2);
        }
        int k = 0;
        for (int pass = 0; pass < 3; ++pass)
            for (int p = 0; p < h->nParts; ++p)
                if (cls[p] == pass) order[k++] = p;
        if (h->kernel == EHYB_KERNEL_PERSISTENT) {
            rc2 = build_cta_tab(h, &v, order);
            if (rc2) { free(order); free(cls); return rc2; }
        }
        free(cls);
        cudaError_t e = cudaMalloc(&h->order, sizeof(int32_t) * (size_t)h->nParts);
        if (e == cudaSuccess) e = cudaMemcpy(h->order, order, sizeof(int32_t) * (size_t)h->nParts, cudaMemcpyHostToDevice);
        free(order);
        if (e != cudaSuccess) return ehyb_fail(EHYB_ERR_CUDA, "dispatch order: %s", cudaGetErrorString(e));
        return EHYB_OK;
    };
    MG_TRY(s, body());
    *out = s;
    return EHYB_OK;
}

extern "C" int ehyb_mg_p2p_export(ehyb_mg_session *s, void *blob)
{
    if (!s || !blob || s->exchange != EHYB_MG_P2P) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_p2p_export: not a peer-memory session");
    CU(cudaSetDevice(s->h->device));
    P2PBlob b;
    memset(&b, 0, sizeof b);
    CU(cudaIpcGetMemHandle(&b.handle, s->shared));
    b.haloStride = (int64_t)s->haloStride;
    b.flagsOffset = (int64_t)(2 * s->haloStride);
    b.nHalo = s->nHalo;
    b.device = s->h->device;
    b.rank = s->rank;
    b.pushCtas = s->pushCtas;
    b.mboxOffset = (int64_t)s->mboxOffset;
    memcpy(blob, &b, sizeof b);
    return EHYB_OK;
}

/* connect `s` to its neighbours described by B[] (one entry per rank); localBase != NULL: the
 * neighbours live in THIS process (ehyb_mg_p2p_connect_local) and localBase[g] is rank g's shared
 * allocation itself, else it is mapped through the CUDA IPC handle of B[g] */
static int p2p_connect_impl(ehyb_mg_session *s, const P2PBlob *B, const int64_t *recvOffsetOnPeer, void *const *localBase)
{
    if (s->connected) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_p2p_connect: already connected");
    CU(cudaSetDevice(s->h->device));
    const int R = s->nranks;
    s->peerBase = (void **)calloc((size_t)R, sizeof(void *));
    s->peerIsLocal = localBase != NULL;
    double **dst[2] = {(double **)malloc(sizeof(double *) * (size_t)(s->nSend ? s->nSend : 1)),
                       (double **)malloc(sizeof(double *) * (size_t)(s->nSend ? s->nSend : 1))};
    uint32_t **flagAddr = (uint32_t **)malloc(sizeof(uint32_t *) * (size_t)R);
    unsigned char **mbox = (unsigned char **)calloc((size_t)R, sizeof(unsigned char *));
    int rc = EHYB_OK;
    auto body = [&]() -> int {
        if (!s->peerBase || !dst[0] || !dst[1] || !flagAddr || !mbox) return ehyb_fail(EHYB_ERR_NOMEM, "p2p connect: out of memory");
        s->recvMask = s->nbrMask = 0;
        s->nPeers = 0;
        int64_t so = 0;
        mbox[s->rank] = s->shared + s->mboxOffset;
        for (int g = 0; g < R; ++g) {
            const int64_t sc = s->sendCount[g], rcv = s->recvCount[g];
            if (g == s->rank) { so += sc; continue; }
            if (B[g].rank != g) return ehyb_fail(EHYB_ERR_ARG, "p2p connect: blob %d describes rank %d", g, B[g].rank);
            if (recvOffsetOnPeer[g] < 0 || recvOffsetOnPeer[g] + sc > B[g].nHalo)
                return ehyb_fail(EHYB_ERR_ARG, "p2p connect: %lld entries at %lld do not fit rank %d's halo of %lld", (long long)sc,
                                 (long long)recvOffsetOnPeer[g], g, (long long)B[g].nHalo);
            const bool sameDevice = B[g].device == s->h->device; /* ranks sharing a GPU ($EHYB_MG_SHARE_DEVICE, tests) */
            int can = sameDevice ? 1 : 0;
            if (!sameDevice) CU(cudaDeviceCanAccessPeer(&can, s->h->device, B[g].device));
            if (!can) return ehyb_fail(EHYB_ERR_PEER, "GPU %d cannot access GPU %d (rank %d) directly", s->h->device, B[g].device, g);
            if (localBase) {
                if (!sameDevice) {
                    cudaError_t e = cudaDeviceEnablePeerAccess(B[g].device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return ehyb_fail(EHYB_ERR_PEER, "cudaDeviceEnablePeerAccess(%d): %s", B[g].device, cudaGetErrorString(e));
                    cudaGetLastError();
                }
                s->peerBase[g] = localBase[g];
            } else {
                cudaError_t e = cudaIpcOpenMemHandle(&s->peerBase[g], B[g].handle, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    s->peerBase[g] = NULL;
                    return ehyb_fail(EHYB_ERR_PEER, "cudaIpcOpenMemHandle(rank %d): %s", g, cudaGetErrorString(e));
                }
            }
            unsigned char *base = (unsigned char *)s->peerBase[g];
            mbox[g] = base + B[g].mboxOffset;
            if (sc == 0 && rcv == 0) continue; /* not a neighbour: mapped for the all-reduce mailbox only */
            for (int64_t k = 0; k < sc; ++k)
                for (int b = 0; b < 2; ++b)
                    dst[b][so + k] = (double *)(base + (size_t)b * (size_t)B[g].haloStride) + recvOffsetOnPeer[g] + k;
            flagAddr[s->nPeers++] = (uint32_t *)(base + B[g].flagsOffset) + (size_t)s->rank * kMaxPushCtas;
            s->nbrMask |= 1u << g;
            if (rcv > 0) s->recvMask |= 1u << g;
            so += sc;
        }
        for (int b = 0; b < 2; ++b) {
            CU(cudaMalloc(&s->pushDst_d[b], sizeof(double *) * (size_t)(s->nSend ? s->nSend : 1)));
            if (s->nSend) CU(cudaMemcpy(s->pushDst_d[b], dst[b], sizeof(double *) * (size_t)s->nSend, cudaMemcpyHostToDevice));
        }
        int32_t *ppc = (int32_t *)malloc(sizeof(int32_t) * (size_t)R);
        if (!ppc) return ehyb_fail(EHYB_ERR_NOMEM, "p2p connect: out of memory");
        for (int g = 0; g < R; ++g) {
            ppc[g] = B[g].pushCtas;
            if (ppc[g] < 1 || ppc[g] > kMaxPushCtas) { free(ppc); return ehyb_fail(EHYB_ERR_ARG, "p2p connect: rank %d announces %d pushing CTAs", g, B[g].pushCtas); }
        }
        cudaError_t e2 = cudaMalloc(&s->peerPushCtas_d, sizeof(int32_t) * (size_t)R);
        if (e2 == cudaSuccess) e2 = cudaMemcpy(s->peerPushCtas_d, ppc, sizeof(int32_t) * (size_t)R, cudaMemcpyHostToDevice);
        free(ppc);
        if (e2 != cudaSuccess) return ehyb_fail(EHYB_ERR_CUDA, "p2p connect: %s", cudaGetErrorString(e2));
        CU(cudaMalloc(&s->peerFlag_d, sizeof(uint32_t *) * (size_t)(s->nPeers ? s->nPeers : 1)));
        if (s->nPeers) CU(cudaMemcpy(s->peerFlag_d, flagAddr, sizeof(uint32_t *) * (size_t)s->nPeers, cudaMemcpyHostToDevice));
        CU(cudaMalloc(&s->mboxPeer_d, sizeof(unsigned char *) * (size_t)R));
        CU(cudaMemcpy(s->mboxPeer_d, mbox, sizeof(unsigned char *) * (size_t)R, cudaMemcpyHostToDevice));
        CU(cudaDeviceSynchronize());
        return EHYB_OK;
    };
    rc = body();
    free(dst[0]); free(dst[1]); free(flagAddr); free(mbox);
    if (rc) return rc;
    s->connected = 1;
    return EHYB_OK;
}

extern "C" int ehyb_mg_p2p_connect(ehyb_mg_session *s, const void *blobs, const int64_t *recvOffsetOnPeer)
{
    if (!s || !blobs || !recvOffsetOnPeer || s->exchange != EHYB_MG_P2P) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_p2p_connect: bad argument");
    return p2p_connect_impl(s, (const P2PBlob *)blobs, recvOffsetOnPeer, NULL);
}

/* All the ranks in ONE process (one host thread per GPU, or one thread driving them all): the
 * sessions reach each other's halo buffers through plain peer access (cudaDeviceEnablePeerAccess)
 * instead of CUDA IPC mappings, which do not work inside the exporting process.  sessions[r] =
 * rank r's session (created with ehyb_mg_session_create_p2p on its own device).  This is what the
 * C driver uses (bin/spmv.out -G N): no launcher, no Python, no torch.distributed. */
extern "C" int ehyb_mg_p2p_connect_local(ehyb_mg_session *const *sessions, int nranks)
{
    if (!sessions || nranks <= 0 || nranks > 32) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_p2p_connect_local: bad argument");
    P2PBlob *B = (P2PBlob *)calloc((size_t)nranks, sizeof(P2PBlob));
    void **base = (void **)calloc((size_t)nranks, sizeof(void *));
    int64_t *off = (int64_t *)calloc((size_t)nranks, sizeof(int64_t));
    if (!B || !base || !off) { free(B); free(base); free(off); return ehyb_fail(EHYB_ERR_NOMEM, "p2p connect: out of memory"); }
    int rc = EHYB_OK;
    for (int g = 0; g < nranks && rc == EHYB_OK; ++g) {
        ehyb_mg_session *s = sessions[g];
        if (!s || s->exchange != EHYB_MG_P2P || s->nranks != nranks || s->rank != g) { rc = ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_p2p_connect_local: sessions[%d] is not rank %d of %d (peer-memory exchange)", g, g, nranks); break; }
        B[g].haloStride = (int64_t)s->haloStride; B[g].flagsOffset = (int64_t)(2 * s->haloStride); B[g].nHalo = s->nHalo;
        B[g].device = s->h->device; B[g].rank = g; B[g].pushCtas = s->pushCtas;
        base[g] = s->shared;
    }
    for (int r = 0; r < nranks && rc == EHYB_OK; ++r) {
        /* rank r's entries start in rank g's halo list behind those of the ranks below r */
        for (int g = 0; g < nranks; ++g) {
            off[g] = 0;
            for (int q = 0; q < r; ++q) off[g] += sessions[g]->recvCount[q];
        }
        rc = p2p_connect_impl(sessions[r], B, off, base);
    }
    free(B); free(base); free(off);
    return rc;
}

/* Scalar all-reduce over peer memory: lane g of one warp stores this rank's `count` partial sums into
 * rank g's mailbox (slot of this rank, parity of the epoch), releases them with the epoch word
 * (st.release.sys), then waits for rank g's contribution in its OWN mailbox (ld.acquire.sys) and the
 * warp adds the contributions in RANK ORDER - the same order on every GPU, so every rank holds the
 * same bits and takes the same decisions from them.  A mailbox slot of parity b is written again two
 * all-reduces later; a rank can only be there after it has received everybody's contribution to the
 * one in between, which every rank sends after it has read this one: two parities are enough.
 * Bounded like every wait on a peer (status word, ehyb_mg_status). */
__global__ void __launch_bounds__(32) ehyb_allreduce_kernel(double *vals, int count, unsigned char *const *mboxPeer, int rank, int nranks,
                                                            uint32_t epoch, unsigned long long timeoutNs, uint32_t *status)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int lane = threadIdx.x;
    const uint32_t b = epoch & 1u;
    const size_t valsOff = (static_cast<size_t>(b) * nranks) * kMboxVals * sizeof(double);
    const size_t flagOff = static_cast<size_t>(2) * nranks * kMboxVals * sizeof(double) + static_cast<size_t>(b) * nranks * sizeof(uint32_t);
    double mine[kMboxVals];
#pragma unroll
    for (int k = 0; k < kMboxVals; ++k) mine[k] = k < count ? __ldcg(vals + k) : 0.0;
    double got[kMboxVals];
#pragma unroll
    for (int k = 0; k < kMboxVals; ++k) got[k] = 0.0;
    bool ok = true;
    if (lane < nranks) {
        unsigned char *peer = mboxPeer[lane];
        double *dst = reinterpret_cast<double *>(peer + valsOff) + static_cast<size_t>(rank) * kMboxVals;
#pragma unroll
        for (int k = 0; k < kMboxVals; ++k) dst[k] = mine[k];
        st_release_sys_u32(reinterpret_cast<uint32_t *>(peer + flagOff) + rank, epoch);
        /* rank `lane`'s contribution, in my own mailbox */
        unsigned char *own = mboxPeer[rank];
        const uint32_t *flag = reinterpret_cast<const uint32_t *>(own + flagOff) + lane;
        const unsigned long long t0 = global_timer_ns();
        unsigned polls = 0;
        while (ld_acquire_sys_u32(flag) != epoch) {
            if ((++polls & 255u) == 0 && timeoutNs && global_timer_ns() - t0 > timeoutNs) { ok = false; break; }
        }
        const double *src = reinterpret_cast<const double *>(own + valsOff) + static_cast<size_t>(lane) * kMboxVals;
#pragma unroll
        for (int k = 0; k < kMboxVals; ++k) got[k] = __ldcg(src + k);
    }
    if (!__all_sync(0xffffffffu, ok)) {
        if (lane == 0) *status = 1u; /* a rank did not show up: the values are left alone, the host reports it */
        return;
    }
#pragma unroll
    for (int k = 0; k < kMboxVals; ++k) {
        double sum = 0.0;
        for (int g = 0; g < nranks; ++g) sum += __shfl_sync(0xffffffffu, got[k], g);
        if (lane == 0 && k < count) vals[k] = sum;
    }
}

/* In-stream all-reduce (sum) of `count` <= 4 doubles at vals_d (device memory, replaced by the sum over
 * the ranks), on the session stream.  Collective: every rank calls it the same number of times.
 * Peer-memory sessions use the mailbox above; NCCL sessions ncclAllReduce on the session stream. */
extern "C" int ehyb_mg_allreduce_sum(ehyb_mg_session *s, double *vals_d, int count)
{
    if (!s || !vals_d || count < 1 || count > kMboxVals) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_allreduce_sum: bad argument (1..%d values)", kMboxVals);
    ehyb_handle *h = s->h;
    CU(cudaSetDevice(h->device));
    if (s->nranks == 1) return EHYB_OK;
    if (s->exchange != EHYB_MG_P2P) {
        NC(g_nccl.AllReduce(vals_d, vals_d, (size_t)count, ncclDouble, ncclSum, s->comm, h->stream));
        return EHYB_OK;
    }
    if (!s->connected) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_allreduce_sum: call ehyb_mg_p2p_connect first");
    s->arEpoch += 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(32);
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = h->pdl ? 1 : 0;
    CU(cudaLaunchKernelEx(&cfg, ehyb_allreduce_kernel, vals_d, count, (unsigned char *const *)s->mboxPeer_d, s->rank, s->nranks, s->arEpoch,
                          s->timeoutNs, s->status_d));
    return EHYB_OK;
}

extern "C" int ehyb_mg_status(ehyb_mg_session *s, int *timed_out)
{
    if (!s || !timed_out) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_status: NULL");
    *timed_out = 0;
    if (s->exchange != EHYB_MG_P2P) return EHYB_OK;
    CU(cudaSetDevice(s->h->device));
    CU(cudaStreamSynchronize(s->h->stream));
    *timed_out = *(volatile uint32_t *)s->status_h != 0;
    return EHYB_OK;
}

extern "C" int ehyb_mg_launches_per_spmv(const ehyb_mg_session *s)
{
    if (!s) return 0;
    const int ovf = s->h->nOvf > 0 ? 1 : 0;
    return s->exchange == EHYB_MG_P2P ? 1 + ovf : 2 + ovf + (s->nSend > 0 ? 1 : 0); /* NCCL: pack + send/recv kernel + main */
}

/* peer-memory product: ONE launch (plus the overflow kernel if the block has overflow
 * entries); the exchange happens inside the main kernel (ehyb_kernels.cuh, PeerArgs) */
static int mg_spmv_p2p(ehyb_mg_session *s, double *x_d, double *y_d, double *dot_d = NULL)
{
    ehyb_handle *h = s->h;
    if (!s->connected) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_spmv: call ehyb_mg_p2p_connect first");
    s->epoch += 1;
    const int b = (int)(s->epoch & 1u);
    PeerArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.xh = (const double *)(s->shared + (size_t)b * s->haloStride);
    pa.flags = (const uint32_t *)(s->shared + 2 * s->haloStride);
    pa.status = s->status_d;
    pa.pushIdx = s->sendIdx_d;
    pa.pushDst = s->pushDst_d[b];
    pa.peerFlag = s->peerFlag_d;
    pa.peerPushCtas = s->peerPushCtas_d;
    pa.timeoutNs = s->timeoutNs;
    pa.epoch = s->epoch;
    pa.recvMask = s->recvMask;
    pa.nbrMask = s->nbrMask;
    pa.nranks = s->nranks;
    pa.pushCount = (int)s->nSend;
    pa.pushCtas = s->pushCtas;
    pa.nPeers = s->nPeers;
    int rc = launch_main(h, x_d, y_d, h->stream, &pa, dot_d);
    return rc ? rc : launch_overflow(h, x_d, y_d, h->stream, &pa);
}

/*
 * One distributed product, NCCL exchange.  x_d holds the local x in its first n entries; the
 * halo part x_d[n, n+nHalo) is filled here.  Order of work:
 *   comm stream : pack (gather the x entries peers need) -> grouped ncclSend/ncclRecv
 *   main stream : main kernel (needs only local x: every halo entry lives in the overflow
 *                 list) || exchange ; then the overflow kernel after the halo has arrived.
 */
static int mg_spmv_nccl(ehyb_mg_session *s, double *x_d, double *y_d)
{
    ehyb_handle *h = s->h;
    CU(cudaEventRecord(s->evX, h->stream)); /* x is ready / previous product done with the halo */
    CU(cudaStreamWaitEvent(s->commStream, s->evX, 0));
    if (s->nSend > 0) {
        ehyb_pack_kernel<<<(unsigned)((s->nSend + 255) / 256), 256, 0, s->commStream>>>(x_d, s->sendIdx_d, s->sendBuf_d, s->nSend);
        CU(cudaGetLastError());
    }
    NC(g_nccl.GroupStart());
    int64_t so = 0, ro = 0;
    for (int g = 0; g < s->nranks; ++g) {
        if (s->sendCount[g] > 0) NC(g_nccl.Send(s->sendBuf_d + so, (size_t)s->sendCount[g], ncclDouble, g, s->comm, s->commStream));
        if (s->recvCount[g] > 0) NC(g_nccl.Recv(x_d + h->n + ro, (size_t)s->recvCount[g], ncclDouble, g, s->comm, s->commStream));
        so += s->sendCount[g];
        ro += s->recvCount[g];
    }
    NC(g_nccl.GroupEnd());
    CU(cudaEventRecord(s->evHalo, s->commStream));
    int rc = launch_main(h, x_d, y_d, h->stream, NULL);
    if (rc) return rc;
    CU(cudaStreamWaitEvent(h->stream, s->evHalo, 0));
    return launch_overflow(h, x_d, y_d, h->stream, NULL);
}

extern "C" int ehyb_mg_spmv(ehyb_mg_session *s, double *x_d, double *y_d)
{
    if (!s || !x_d || !y_d) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_spmv: NULL argument");
    CU(cudaSetDevice(s->h->device));
    return s->exchange == EHYB_MG_P2P ? mg_spmv_p2p(s, x_d, y_d) : mg_spmv_nccl(s, x_d, y_d);
}

/* The distributed product with the fused dot (ehyb_spmv_dot): *dot_d += x_own . y over this rank's
 * rows; the sum over the ranks is the caller's (ehyb_mg_allreduce_sum).  Peer-memory sessions whose
 * block has no overflow entries only (ehyb_mg_spmv_dot_supported). */
extern "C" int ehyb_mg_spmv_dot_supported(const ehyb_mg_session *s)
{
    return s && s->exchange == EHYB_MG_P2P && ehyb_spmv_dot_supported(s->h);
}

extern "C" int ehyb_mg_spmv_dot(ehyb_mg_session *s, double *x_d, double *y_d, double *dot_d)
{
    if (!s || !x_d || !y_d || !dot_d) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_spmv_dot: NULL argument");
    if (!ehyb_mg_spmv_dot_supported(s)) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_spmv_dot: not a peer-memory session with a single-launch product");
    CU(cudaSetDevice(s->h->device));
    return mg_spmv_p2p(s, x_d, y_d, dot_d);
}

/* device of a session (a solver allocates its vectors there) */
extern "C" int ehyb_session_device(const ehyb_handle *h) { return h ? h->device : -1; }
extern "C" int ehyb_mg_session_ranks(const ehyb_mg_session *s, int *rank, int *nranks)
{
    if (!s) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_session_ranks: NULL");
    if (rank) *rank = s->rank;
    if (nranks) *nranks = s->nranks;
    return EHYB_OK;
}

/* Pipelined stream of distributed products with HOST vectors: for i in [0, count): y_h[i] =
 * A_block [x_h[i] | halo], x_h[i] / y_h[i] the n local entries in (pinned) host memory.  As
 * ehyb_spmv_host_batch: the H2D of product i+1 and the D2H of product i-1 overlap the kernel of
 * product i (two buffer pairs, three streams); the halo travels between the GPUs inside the
 * products.  Collective: every rank calls it with the same count. */
extern "C" int ehyb_mg_spmv_host_batch(ehyb_mg_session *s, const double *const *x_h, double *const *y_h, int count)
{
    if (!s || !x_h || !y_h || count < 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_spmv_host_batch: bad argument");
    ehyb_handle *h = s->h;
    CU(cudaSetDevice(h->device));
    int rc = ensure_pipeline(h);
    if (rc) return rc;
    for (int i = 0; i < count; ++i) {
        const int b = i & 1;
        if (i >= 2) CU(cudaStreamWaitEvent(h->h2d, h->evK[b], 0));
        CU(cudaMemcpyAsync(h->xb[b], x_h[i], sizeof(double) * (size_t)h->n, cudaMemcpyHostToDevice, h->h2d));
        CU(cudaEventRecord(h->evX[b], h->h2d));
        CU(cudaStreamWaitEvent(h->stream, h->evX[b], 0));
        if (i >= 2) CU(cudaStreamWaitEvent(h->stream, h->evY[b], 0));
        rc = s->exchange == EHYB_MG_P2P ? mg_spmv_p2p(s, h->xb[b], h->yb[b]) : mg_spmv_nccl(s, h->xb[b], h->yb[b]);
        if (rc) return rc;
        CU(cudaEventRecord(h->evK[b], h->stream));
        CU(cudaStreamWaitEvent(h->d2h, h->evK[b], 0));
        CU(cudaMemcpyAsync(y_h[i], h->yb[b], sizeof(double) * (size_t)h->n, cudaMemcpyDeviceToHost, h->d2h));
        CU(cudaEventRecord(h->evY[b], h->d2h));
    }
    CU(cudaStreamSynchronize(h->h2d));
    CU(cudaStreamSynchronize(h->stream));
    if (s->commStream) CU(cudaStreamSynchronize(s->commStream));
    CU(cudaStreamSynchronize(h->d2h));
    return peer_check(h);
}

extern "C" int ehyb_mg_session_handle(ehyb_mg_session *s, ehyb_handle **h)
{
    if (!s || !h) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_session_handle: NULL");
    *h = s->h;
    return EHYB_OK;
}

/* `iters` distributed products of the session's own x between two events on the main
 * stream (the caller brackets this with its barrier and takes the max over ranks). */
extern "C" int ehyb_mg_time_spmv(ehyb_mg_session *s, int warmup, int iters, float *ms_total)
{
    if (!s || !ms_total || iters <= 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_time_spmv: bad argument");
    ehyb_handle *h = s->h;
    CU(cudaSetDevice(h->device));
    for (int i = 0; i < warmup; ++i) {
        int rc = ehyb_mg_spmv(s, h->x, h->y);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(h->stream));
    if (s->commStream) CU(cudaStreamSynchronize(s->commStream));
    CU(cudaEventRecord(h->ev0, h->stream));
    for (int i = 0; i < iters; ++i) {
        int rc = ehyb_mg_spmv(s, h->x, h->y);
        if (rc) return rc;
    }
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (s->commStream) CU(cudaStreamSynchronize(s->commStream));
    CU(cudaEventElapsedTime(ms_total, h->ev0, h->ev1));
    int timedOut = 0;
    int rc = ehyb_mg_status(s, &timedOut);
    if (rc) return rc;
    if (timedOut) return ehyb_fail(EHYB_ERR_PEER, "a neighbour did not deliver its halo within the time limit");
    return EHYB_OK;
}
