/*
 * session.c -- the reference's session entry points on top of the C ABI.
 *
 * spmvGPuEHYB mirrors reference spmv.cu:61-133 (convert -> upload -> 10 warm-ups -> MAXIter
 * timed products of the same x -> download y -> report) and matrixVectorEHYB[_small] mirror
 * the per-product launchers (kernel.cu:490-552).  Differences, all deliberate:
 *   - the timed loop is measured with CUDA events on the session stream, and reported both that
 *     way and the reference's way (wall clock around one H2D of x, the products and one D2H of
 *     y, spmv.cu:108-119);
 *   - every product does all of its work (the reference's remainder phase only runs in the
 *     first launch), and the flop count in the report is still 2*nnz;
 *   - errors abort with a message instead of being ignored; everything is freed.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "common.h"
#include "kernel.h"


#define EHYB_NOMINAL_HBM_GBS 8000.0 /* BASELINE.json: "~8 TB/s per-GPU roofline" */

static double measured_hbm_gbs(void)
{
    const char *s = getenv("EHYB_MEASURED_HBM_GBS");
    return s && s[0] ? atof(s) : 0.0;
}

static void session_on_layout(const ehyb_layout *L, const double *vectorIn, double *vectorOut, const int MAXIter, int *realIter);

void spmvGPuEHYB(matrixCOO *localMatrix, const double *vectorIn, double *vectorOut, const int MAXIter, int *realIter)
{
    if (!localMatrix || !vectorIn || !vectorOut) {
        ehyb_fail(EHYB_ERR_ARG, "NULL argument");
        ehyb_die("spmvGPuEHYB");
    }
    ehyb_layout *L = NULL;
    ehyb_layout_opts lo;
    memset(&lo, 0, sizeof lo);
    lo.er_fill = -1.0; /* automatic */
    const char *fillEnv = getenv("EHYB_ER_FILL");
    if (fillEnv && fillEnv[0]) lo.er_fill = atof(fillEnv);
    if (ehyb_layout_build(localMatrix, &lo, &L)) ehyb_die("spmvGPuEHYB: format build");
    session_on_layout(L, vectorIn, vectorOut, MAXIter, realIter);
    ehyb_layout_free(L);
}

/* The same session on a layout that already exists (built by the caller, or loaded from the
 * binary cache, cache.c): everything of spmvGPuEHYB after the format build. */
void spmvGPuEHYB_layout(const ehyb_layout *L, const double *vectorIn, double *vectorOut, const int MAXIter, int *realIter)
{
    if (!L || !vectorIn || !vectorOut) {
        ehyb_fail(EHYB_ERR_ARG, "NULL argument");
        ehyb_die("spmvGPuEHYB_layout");
    }
    session_on_layout(L, vectorIn, vectorOut, MAXIter, realIter);
}

static void session_on_layout(const ehyb_layout *L, const double *vectorIn, double *vectorOut, const int MAXIter, int *realIter)
{
    ehyb_handle *h = NULL;
    ehyb_layout_view v;
    ehyb_layout_get(L, &v);
    const int64_t totalNum = v.nnz;
    /* the reference's converter lines (convert.c:140, :310; spmv.cu:82), same meaning */
    printf("toER is %lld, kernel calculation is %lld\n", (long long)(v.nnz - v.nnzEll), (long long)v.nnzEll);
    printf("wasteElement is %lld\n", (long long)v.padEll);
    printf("sizeER is %lld\n", (long long)(v.nnzRemInSlice + v.padRem + v.nOverflow));

    ehyb_session_opts so;
    ehyb_session_opts_default(&so);
    const char *devEnv = getenv("EHYB_DEVICE");
    if (devEnv && devEnv[0]) so.device = atoi(devEnv);
    if (ehyb_upload(L, &so, &h)) ehyb_die("spmvGPuEHYB: upload");
    int threads = 0, ctas = 0, grid = 0, persist = 0;
    int64_t smem = 0;
    ehyb_session_info(h, &threads, &ctas, &grid, &smem, &persist);
    printf("EHYB-B200: %d partitions x %d CTA, window %d (%lld B smem), %d threads/CTA, %d CTA/SM, %d slices, "
           "overflow %lld entries, L2 window on x %s, kernel %s (grid %d)\n",
           v.nParts, v.ctasPerPart, v.W, (long long)smem, threads, ctas, v.nSlices, (long long)v.nOverflow,
           persist ? "on" : "off", ehyb_session_kernel(h), grid);

    if (ehyb_set_x(h, vectorIn)) ehyb_die("spmvGPuEHYB: H2D");
    float ms = 0.f, kms = 0.f;
    /* warm-up: 10 products (spmv.cu:100-106); then the event-timed loop */
    if (ehyb_time_spmv(h, 10, MAXIter > 0 ? MAXIter : 1, &ms, &kms)) ehyb_die("spmvGPuEHYB: timed loop");
    if (ehyb_get_y(h, vectorOut)) ehyb_die("spmvGPuEHYB: D2H");

    /* the reference's own number: wall clock around H2D x + MAXIter products + D2H y */
    struct timeval t0, t1;
    double *xd = NULL, *yd = NULL;
    ehyb_session_vectors(h, &xd, &yd);
    gettimeofday(&t0, NULL);
    if (ehyb_set_x(h, vectorIn)) ehyb_die("spmvGPuEHYB: H2D");
    int iter = 0;
    while (iter < MAXIter) {
        if (ehyb_spmv(h, xd, yd)) ehyb_die("spmvGPuEHYB: product");
        iter++;
    }
    if (ehyb_get_y(h, vectorOut)) ehyb_die("spmvGPuEHYB: D2H");
    gettimeofday(&t1, NULL);
    const double wallMs = ((double)(t1.tv_sec - t0.tv_sec) * 1e6 + (double)(t1.tv_usec - t0.tv_usec)) / 1000.0;
    printf("iter is %d, time is %f ms, GPU Gflops is %f\n ", iter, wallMs, (1e-9 * (2.0 * totalNum) * 1000 * iter) / wallMs);

    const int it = MAXIter > 0 ? MAXIter : 1;
    const double perIterMs = ms / it, perKernelMs = kms / it;
    const double gbs = (double)v.algBytes / (perIterMs * 1e6), kgbs = (double)v.algBytes / (perKernelMs * 1e6);
    printf("\nEHYB-B200 events: %.3f us per product, %.1f GFLOP/s, %.1f GB/s algorithmic (%.1f %% of nominal %.0f GB/s",
           perIterMs * 1e3, 2.0 * totalNum / (perIterMs * 1e6), gbs, 100.0 * gbs / EHYB_NOMINAL_HBM_GBS, EHYB_NOMINAL_HBM_GBS);
    if (measured_hbm_gbs() > 0) printf(", %.1f %% of measured %.0f GB/s", 100.0 * gbs / measured_hbm_gbs(), measured_hbm_gbs());
    printf("); main kernel alone %.3f us, %.1f GB/s\n", perKernelMs * 1e3, kgbs);
    printf("EHYB-B200 bytes: algorithmic %lld (CSR-equivalent %lld), format stored %lld\n", (long long)v.algBytes,
           (long long)(12LL * v.nnz + 4 * (v.n + 1) + 16 * v.n), (long long)v.formatBytes);
    if (realIter) *realIter = iter;
    ehyb_free(h);
}

static ehyb_handle *session_of(matrixEHYB *m, const char *who)
{
    if (!m || !m->b200) {
        ehyb_fail(EHYB_ERR_ARG, "the matrixEHYB does not describe an uploaded session (use ehyb_upload + ehyb_describe)");
        ehyb_die(who);
    }
    return (ehyb_handle *)m->b200;
}

void matrixVectorEHYB(matrixEHYB *inputMatrix, double *vector_in_d, double *vector_out_d)
{
    if (ehyb_spmv(session_of(inputMatrix, "matrixVectorEHYB"), vector_in_d, vector_out_d)) ehyb_die("matrixVectorEHYB");
}

void matrixVectorEHYB_small(matrixEHYB *inputMatrix_d, int *biasIdxBlock_d, double *vectorIn_d, double *vectorOut_d)
{
    (void)biasIdxBlock_d; /* the reference's per-partition work counter (kernel.cu:232,254): not needed */
    if (ehyb_spmv(session_of(inputMatrix_d, "matrixVectorEHYB_small"), vectorIn_d, vectorOut_d)) ehyb_die("matrixVectorEHYB_small");
}
