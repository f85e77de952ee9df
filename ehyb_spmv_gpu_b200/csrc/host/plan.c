/*
 * plan.c -- partition-parameter choice.
 *
 * The reference picks (nParts, vectorCacheSize, kernelPerPart) in the reader from
 * compile-time constants of an 82-SM / 93 KB device (solver_test.c:158-182, kernel.h:20-25)
 * and stores the window in a 16-bit integer, which wraps for n > ~2.6 M (SURVEY.md B-7,
 * Appendix C).  ehyb_plan() derives the parameters from the queried device instead:
 *
 *   - one CTA is resident per SM and its shared memory is split between the x window, the
 *     remainder cache and the staging slots of the matrix stream (the staged kernel keeps
 *     ~130 KB of TMA copies in flight per SM); the window may use what is left;
 *   - the number of partitions is a multiple of the SM count (two per SM by default), so
 *     every SM streams the same number of partitions;
 *   - the window is sized from the partition size with 2.5 % head-room for the partitioner's
 *     imbalance (mt-metis: <= 0.1 % on stencils, 3 % on the elasticity graph)
 *     - rows beyond the window still work, they just have no ELL entries;
 *   - matrices too small to give every SM `minRows` rows get fewer partitions and several
 *     CTAs per partition (the reference's "_small" idea).
 *
 * Tunables can be overridden through the environment for experiments:
 * EHYB_PARTS_PER_SM, EHYB_CTAS_PER_SM, EHYB_THREADS, EHYB_CTAS_PER_PART.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "common.h"

void ehyb_device_info_b200(ehyb_device_info *d)
{
    memset(d, 0, sizeof *d);
    d->device = -1;
    d->sm_count = 148;
    d->smem_optin_bytes = 232448;  /* 227 KB */
    d->smem_per_sm_bytes = 233472; /* 228 KB */
    d->l2_bytes = 126 * 1024 * 1024;
    d->max_persist_l2_bytes = 79 * 1024 * 1024;
    d->cc_major = 10;
    d->cc_minor = 0;
    d->hbm_bytes = (size_t)183359 * 1024 * 1024;
    strcpy(d->name, "NVIDIA B200 (nominal)");
}

static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return s && s[0] ? atoi(s) : dflt;
}

#define EHYB_SMEM_RESERVE 512 /* mbarriers in front of the window */

int ehyb_plan(int n, const ehyb_device_info *dev, ehyb_plan_t *out) { return ehyb_plan_kernel(n, dev, EHYB_KERNEL_STAGED, out); }

/*
 * kernel = EHYB_KERNEL_PERSISTENT: partitions sized for the persistent kernel, which holds TWO
 * {window, remainder cache} buffers next to its staging slots and hides the partition start-up:
 * three partitions per SM instead of two (config 2: 92.2 us vs 94.3 us staged) and a window of at
 * most ~6 K entries, so that two buffers leave room for >= 16 warps of staging (256^3: 805 us vs
 * 840 us).  Matrices too small for three partitions of minRows rows per SM get the staged plan;
 * ehyb_upload falls back to the staged kernel by itself when the persistent one does not fit
 * (very long remainder-cache lists).
 */
int ehyb_plan_kernel(int n, const ehyb_device_info *dev, int kernel, ehyb_plan_t *out)
{
    if (n <= 0 || !dev || !out || dev->sm_count <= 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_plan: bad argument");
    const int sms = dev->sm_count;
    const int persistent = kernel == EHYB_KERNEL_PERSISTENT && (double)n / (3.0 * sms) >= 2048.0 && !getenv("EHYB_PARTS_PER_SM");
    int partsPerSM = env_int("EHYB_PARTS_PER_SM", persistent ? 3 : 2);
    int ctasPerSM = env_int("EHYB_CTAS_PER_SM", 1);
    int threads = env_int("EHYB_THREADS", 0);
    /* shared memory next to the window: staging slots of the matrix stream (what is in flight
     * per SM) and the remainder cache (EHYB_DEFAULT_CACHE_CAP doubles) */
    const long staging = (long)env_int("EHYB_STAGING_KB", 128) * 1024 + (long)EHYB_DEFAULT_CACHE_CAP * 8;
    if (partsPerSM < 1) partsPerSM = 1;
    if (ctasPerSM < 1) ctasPerSM = 1;
    if (threads < 0 || threads > 1024 || threads % 32) threads = 0; /* 0: the session decides */
    const int minRows = 2048; /* below this a partition's window no longer amortises its load */

    /* largest window one CTA can hold with ctasPerSM CTAs resident */
    long budget = dev->smem_per_sm_bytes / ctasPerSM - 1024;
    if (budget > dev->smem_optin_bytes) budget = dev->smem_optin_bytes;
    budget -= staging;
    int wMax = (int)((budget - EHYB_SMEM_RESERVE) / 8 - 2);
    if (persistent) {
        /* two buffers of window + ~24 KB of cache next to 16 warps x 2 slots of 2.5 KB and the header */
        const long two = (long)dev->smem_optin_bytes - 1664 - 16L * 2 * 2560;
        const int wP = (int)((two / 2 - 24 * 1024) / 8 - 2);
        if (wP < wMax) wMax = wP;
    }
    wMax -= wMax % 64;
    if (wMax > 65472) wMax = 65472; /* the boundary struct carries W in a uint16_t (spmv.h): 65536 would wrap to 0 */
    if (wMax < 64) return ehyb_fail(EHYB_ERR_ARG, "ehyb_plan: no shared memory for a window");

    int P = sms * partsPerSM, kpp = 1;
    if ((double)n / P < minRows) {
        /* small matrix: fewer partitions, several CTAs each, still ~ctasPerSM CTAs per SM */
        P = (n + minRows - 1) / minRows;
        if (P < 1) P = 1;
        kpp = (sms * ctasPerSM + P - 1) / P;
        if (kpp < 1) kpp = 1;
        if (kpp > 32) kpp = 32;
    }
    /* grow P (in multiples of the SM count) until the partitions fit the window */
    while (ceil((double)n / P * 1.025) > wMax) P += sms;
    int W = (int)ceil((double)n / P * 1.025);
    W = (W + 63) / 64 * 64;
    if (W > wMax) W = wMax;
    kpp = env_int("EHYB_CTAS_PER_PART", kpp);
    if (kpp < 1) kpp = 1;
    out->nParts = P;
    out->W = W;
    out->ctasPerPart = kpp;
    out->threads = threads;
    out->ctasPerSM = ctasPerSM;
    return EHYB_OK;
}

/*
 * The plan a driver should use for a matrix of n rows and nnz entries, and the kernel it is for:
 *   - L2-RESIDENT matrices (matrix data + vectors <= 3/4 of L2, the bound below which the session leaves the L2 hints off; BASELINE config 1: 69 MB): ONE
 *     partition per SM for the staged kernel - all CTAs are resident at once, a product is one wave
 *     and launches chain through programmatic dependent launch.  Measured on 5-point 1024^2
 *     (profiles/r2_sweep_config1_*.log): 13.0 us per product, against 16.7 us with the persistent plan
 *     (3 partitions per CTA: its start-up is the whole cost when the stream comes out of L2), 16.4 us
 *     with the direct kernel, 22.5 us with 2 partitions per SM;
 *   - up to ~40 entries per row: the persistent kernel's plan (3 small partitions per SM);
 *   - denser rows (3-dof elasticity: remainder-cache lists too long for two buffers): the staged plan.
 */
int ehyb_plan_auto(int n, int64_t nnz, const ehyb_device_info *dev, ehyb_plan_t *out, int *kernel)
{
    if (n <= 0 || nnz < 0 || !dev || !out || dev->sm_count <= 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_plan_auto: bad argument");
    const int64_t bytes = 10 * nnz + 16 * (int64_t)n;
    int k = nnz <= 40 * (int64_t)n ? EHYB_KERNEL_PERSISTENT : EHYB_KERNEL_STAGED;
    if (dev->l2_bytes > 0 && bytes <= (int64_t)dev->l2_bytes / 4 * 3 && n >= 64 * dev->sm_count && !getenv("EHYB_PARTS_PER_SM")) {
        /* one partition per SM, if its window fits next to the staging slots */
        long budget = dev->smem_per_sm_bytes - 1024;
        if (budget > dev->smem_optin_bytes) budget = dev->smem_optin_bytes;
        budget -= 128L * 1024 + (long)EHYB_DEFAULT_CACHE_CAP * 8 + EHYB_SMEM_RESERVE;
        int wMax = (int)(budget / 8 - 2);
        wMax -= wMax % 64;
        int W = (int)ceil((double)n / dev->sm_count * 1.025);
        W = (W + 63) / 64 * 64;
        if (W <= wMax && W <= 65472) {
            out->nParts = dev->sm_count; out->W = W; out->ctasPerPart = 1; out->threads = 0; out->ctasPerSM = 1;
            if (kernel) *kernel = EHYB_KERNEL_STAGED;
            return EHYB_OK;
        }
    }
    if (kernel) *kernel = k;
    return ehyb_plan_kernel(n, dev, k, out);
}

int ehyb_plan_reference(int n, int symmetric, ehyb_plan_t *out)
{
    if (n <= 0 || !out) return ehyb_fail(EHYB_ERR_ARG, "ehyb_plan_reference: bad argument");
    /* kernel.h:21-25 */
    const int sm = 82, sm2 = 80, tpb = 1024;
    const size_t maxShared = 93 * 1024;
    int factor = 1, kpp = 0;
    /* the window lives in an int16_t (solver_test.c:55,160): out-of-range doubles wrap */
#define WRAP16(d) ((int16_t)(int32_t)(d))
    int16_t w = WRAP16(ceil((double)n / ((double)factor * sm * tpb)) * tpb);
    if ((size_t)(long)w < maxShared / (2 * sizeof(double))) {
        static const int ks[4] = {8, 5, 4, 2};
        int i = 0;
        do {
            kpp = ks[i++];
            w = WRAP16(kpp * ceil((double)n / ((double)sm2 * tpb)) * tpb);
        } while ((size_t)(long)w * sizeof(double) > maxShared && i < 4);
        out->nParts = (symmetric ? sm2 : sm) / kpp; /* solver_test.c:173 vs :68 */
    } else {
        while ((size_t)(long)w * sizeof(double) > maxShared) {
            factor += 1;
            w = WRAP16(ceil((double)n / ((double)factor * sm * tpb)) * tpb);
        }
        out->nParts = factor * sm;
    }
#undef WRAP16
    out->W = (int)(uint16_t)w;
    out->ctasPerPart = kpp;
    out->threads = tpb;
    out->ctasPerSM = 1;
    return EHYB_OK;
}
