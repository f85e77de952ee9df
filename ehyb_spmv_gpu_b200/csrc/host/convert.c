/*
 * convert.c -- COO -> EHYB in the REFERENCE layout.
 *
 * Byte-for-byte the arrays of reference convert.c:316-369 on every input the reference
 * handles, built in a different way: one pass classifies the rows (in-window count is
 * numInRow2 from the reorder stage), the global remainder ranking is a stable counting
 * sort on the spill count (the reference qsorts all n rows with a comparator that is 0 on
 * ties), and slices are filled in parallel.  The device never consumes this layout: it is
 * the drop-in COO2EHYB and the anchor the Blackwell layout (layout.c) is checked against.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.h"
#include "convert.h"
#include "kernel.h"

static int convert_impl(const matrixCOO *in, matrixEHYB *out, int *sizeBlockELL, int *sizeER, int quiet)
{
    if (!in || !out || !sizeBlockELL || !sizeER) return ehyb_fail(EHYB_ERR_ARG, "COO2EHYB: NULL argument");
    const int n = in->dimension, P = in->nParts, W = in->vectorCacheSize;
    if (n <= 0 || P <= 0 || W <= 0 || W % EHYB_WARP || W > 32767)
        return ehyb_fail(EHYB_ERR_ARG, "COO2EHYB: needs nParts > 0 and a window that is a multiple of 32 below 32768 "
                                       "(got nParts %d, window %d); larger windows exist only in the tuned layout", P, W);
    const int S = W / EHYB_WARP; /* slices per partition: fixed, W/32 (convert.c:80) */
    const int *pb = in->partBoundary, *len = in->numInRow, *inw = in->numInRow2, *rowIdx = in->rowIdx;
    memset(out, 0, sizeof *out);
    out->dimension = n;
    out->nParts = P;
    out->vectorCacheSize = (int16_t)W;
    out->kernelPerPart = in->kernelPerPart;
    out->partBoundary = in->partBoundary; /* aliased, convert.c:330 */

    int rc = EHYB_OK;
    int *spill = (int *)calloc((size_t)n, sizeof(int));      /* numInRowER */
    int *firstReg = (int *)malloc((size_t)P * sizeof(int));  /* first non-long row of a partition */
    int *rankER = (int *)malloc((size_t)n * sizeof(int));    /* reorderListER */
    int16_t *wE = (int16_t *)calloc((size_t)P * S, sizeof(int16_t));
    int *bE = (int *)calloc((size_t)P * S, sizeof(int));
    int *longRows = NULL, *cnt = NULL;
    if (!spill || !firstReg || !rankER || !wE || !bE) { rc = ehyb_fail(EHYB_ERR_NOMEM, "COO2EHYB: out of memory"); goto fail; }

    /* pass 1: long rows, slice widths, spill counts (convert.c:87-135) */
    int nLong = 0;
    for (int p = 0; p < P; ++p) {
        int r = pb[p];
        while (r < pb[p + 1] && inw[r] > EHYB_REF_LONG_ROW) ++r;
        firstReg[p] = r;
        nLong += r - pb[p];
    }
    long long toER = 0;
    int nRowER = 0, maxSpill = 0;
#pragma omp parallel for schedule(static) reduction(+ : toER, nRowER) reduction(max : maxSpill)
    for (int p = 0; p < P; ++p) {
        const int ps = pb[p], pe = pb[p + 1], winEnd = ps + W;
        for (int r = firstReg[p]; r < pe; ++r) {
            int sp;
            if (r < winEnd) {
                const int s = (r - ps) / EHYB_WARP;
                if (inw[r] > wE[p * S + s]) wE[p * S + s] = (int16_t)inw[r];
                sp = len[r] - inw[r];
                if (sp == 0) continue;
            } else {
                sp = len[r]; /* beyond the window: the whole row is remainder (convert.c:128-134) */
            }
            spill[r] = sp;
            toER += sp;
            nRowER += 1; /* counted even when sp == 0 beyond the window, as the reference does */
            if (sp > maxSpill) maxSpill = sp;
        }
    }
    if (!quiet) printf("toER is %lld, kernel calculation is %lld\n", toER, (long long)in->totalNum - toER); /* convert.c:140 */

    long long total = 0;
    for (int b = 0; b < P * S; ++b) { /* convert.c:336-340 */
        bE[b] = (int)total;
        total += (long long)EHYB_WARP * wE[b];
    }
    if (total > INT_MAX) { rc = ehyb_fail(EHYB_ERR_LIMIT, "reference layout: ELL part exceeds 2^31 elements"); goto fail; }
    *sizeBlockELL = (int)total;

    /* remainder ranking: all rows by spill count descending, row ascending (convert.c:8-31,152) */
    cnt = (int *)calloc((size_t)maxSpill + 2, sizeof(int));
    if (!cnt) { rc = ehyb_fail(EHYB_ERR_NOMEM, "COO2EHYB: out of memory"); goto fail; }
    for (int r = 0; r < n; ++r) cnt[maxSpill - spill[r] + 1] += 1;
    for (int k = 0; k <= maxSpill; ++k) cnt[k + 1] += cnt[k];
    for (int r = 0; r < n; ++r) rankER[r] = cnt[maxSpill - spill[r]]++;

    const int nbER = (nRowER + EHYB_WARP - 1) / EHYB_WARP;
    out->numOfRowER = nRowER;
    out->rowVecER = (int *)calloc((size_t)(nRowER ? nRowER : 1), sizeof(int));
    out->widthVecER = (int16_t *)calloc((size_t)(nbER ? nbER : 1), sizeof(int16_t));
    out->biasVecER = (int *)calloc((size_t)(nbER ? nbER : 1), sizeof(int));
    if (!out->rowVecER || !out->widthVecER || !out->biasVecER) { rc = ehyb_fail(EHYB_ERR_NOMEM, "COO2EHYB: out of memory"); goto fail; }
    for (int r = 0; r < n; ++r) { /* convert.c:154-166 */
        if (spill[r] > 0) {
            if (spill[r] > 32767) { rc = ehyb_fail(EHYB_ERR_LIMIT, "reference layout: row %d spills %d entries (int16 width)", r, spill[r]); goto fail; }
            out->rowVecER[rankER[r]] = r;
            int16_t *w = &out->widthVecER[rankER[r] / EHYB_WARP];
            if (spill[r] > *w) *w = (int16_t)spill[r];
        }
    }
    total = 0;
    for (int b = 0; b < nbER; ++b) { /* convert.c:348-354 */
        out->biasVecER[b] = (int)total;
        total += (long long)EHYB_WARP * out->widthVecER[b];
    }
    if (total > INT_MAX) { rc = ehyb_fail(EHYB_ERR_LIMIT, "reference layout: remainder exceeds 2^31 elements"); goto fail; }
    *sizeER = (int)total;

    out->valBlockELL = (double *)calloc((size_t)(*sizeBlockELL ? *sizeBlockELL : 1), sizeof(double));
    out->colBlockELL = (int16_t *)calloc((size_t)(*sizeBlockELL ? *sizeBlockELL : 1), sizeof(int16_t));
    out->valER = (double *)calloc((size_t)(*sizeER ? *sizeER : 1), sizeof(double));
    out->colER = (int *)calloc((size_t)(*sizeER ? *sizeER : 1), sizeof(int));
    if (!out->valBlockELL || !out->colBlockELL || !out->valER || !out->colER) { rc = ehyb_fail(EHYB_ERR_NOMEM, "COO2EHYB: out of memory"); goto fail; }

    /* pass 2: fill (convert.c:207-308).  Every row writes only its own lane of its own
     * slice and its own lane of its remainder slice, so rows are independent. */
    long long waste = 0;
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : waste) reduction(| : bad)
    for (int p = 0; p < P; ++p) {
        const int ps = pb[p], pe = pb[p + 1], winEnd = ps + W;
        for (int s = 0; s < S; ++s) {
            const int w = wE[p * S + s];
            int used = 0;
            for (int l = 0; l < EHYB_WARP; ++l) {
                const int r = ps + s * EHYB_WARP + l;
                if (r >= firstReg[p] && r < pe) used += inw[r];
            }
            waste += (long long)w * EHYB_WARP - used; /* zero padding, convert.c:269-281 */
        }
        for (int r = firstReg[p]; r < pe; ++r) {
            const int inWindowRow = r < winEnd;
            const int lane = (r - ps) % EHYB_WARP;
            const int bias = inWindowRow ? bE[p * S + (r - ps) / EHYB_WARP] : 0;
            const int q = rankER[r];
            const int biasR = spill[r] > 0 ? out->biasVecER[q / EHYB_WARP] + q % EHYB_WARP : 0;
            int kE = 0, kR = 0;
            for (int e = rowIdx[r]; e < rowIdx[r + 1]; ++e) {
                const int c = in->J[e];
                if (in->I[e] != r) bad = 1; /* convert.c:243-246 */
                if (inWindowRow && c >= ps && c < winEnd) {
                    out->colBlockELL[bias + lane + kE * EHYB_WARP] = (int16_t)(c - ps);
                    out->valBlockELL[bias + lane + kE * EHYB_WARP] = in->V[e];
                    ++kE;
                } else {
                    if (kR >= spill[r]) { bad = 1; break; }
                    out->colER[biasR + kR * EHYB_WARP] = c;
                    out->valER[biasR + kR * EHYB_WARP] = in->V[e];
                    ++kR;
                }
            }
            if (inWindowRow && kE != inw[r]) bad = 1; /* numInRow2 must match the entries */
        }
    }
    if (bad) { rc = ehyb_fail(EHYB_ERR_ARG, "COO2EHYB: numInRow/numInRow2/rowIdx do not match the entries"); goto fail; }
    if (!quiet) printf("wasteElement is %lld\n", waste); /* convert.c:310 */

    /* long rows, the reference's intent (convert.c:33-59, 92-101; broken there, B-3) */
    out->nLongVec = nLong;
    if (nLong > 0) {
        longRows = (int *)malloc((size_t)nLong * sizeof(int));
        out->longVecBoundary = (int *)malloc(((size_t)nLong + 1) * sizeof(int));
        if (!longRows || !out->longVecBoundary) { rc = ehyb_fail(EHYB_ERR_NOMEM, "COO2EHYB: out of memory"); goto fail; }
        int k = 0;
        for (int p = 0; p < P; ++p)
            for (int r = pb[p]; r < firstReg[p]; ++r) longRows[k++] = r;
        out->longVecBoundary[0] = 0;
        for (k = 0; k < nLong; ++k) out->longVecBoundary[k + 1] = out->longVecBoundary[k] + len[longRows[k]];
        const int tot = out->longVecBoundary[nLong];
        out->longVecRow = longRows;
        longRows = NULL;
        out->longVecCol = (int *)malloc((size_t)(tot ? tot : 1) * sizeof(int));
        out->longVecVal = (double *)malloc((size_t)(tot ? tot : 1) * sizeof(double));
        if (!out->longVecCol || !out->longVecVal) { rc = ehyb_fail(EHYB_ERR_NOMEM, "COO2EHYB: out of memory"); goto fail; }
        for (k = 0; k < nLong; ++k) {
            const int r = out->longVecRow[k];
            memcpy(out->longVecCol + out->longVecBoundary[k], in->J + rowIdx[r], (size_t)len[r] * sizeof(int));
            memcpy(out->longVecVal + out->longVecBoundary[k], in->V + rowIdx[r], (size_t)len[r] * sizeof(double));
        }
    }
    out->widthVecBlockELL = wE;
    out->biasVecBlockELL = bE;
    out->reorderListER = rankER;
    out->reorderList = in->reorderList;
    free(spill); free(firstReg); free(cnt);
    return EHYB_OK;

fail:
    free(spill); free(firstReg); free(rankER); free(wE); free(bE); free(cnt); free(longRows);
    out->partBoundary = NULL;
    EHYBfreeHost(out);
    return rc;
}

int ehyb_convert_reference_layout(const matrixCOO *in, matrixEHYB *out, int *sizeBlockELL, int *sizeER, int quiet)
{
    return convert_impl(in, out, sizeBlockELL, sizeER, quiet);
}

void COO2EHYB(matrixCOO *in, matrixEHYB *out, int *sizeBlockELL, int *sizeER)
{
    if (convert_impl(in, out, sizeBlockELL, sizeER, 0)) ehyb_die("COO2EHYB");
}

void EHYBfreeHost(matrixEHYB *m)
{
    if (!m) return;
    free(m->reorderListER); free(m->widthVecBlockELL); free(m->biasVecBlockELL);
    free(m->colBlockELL); free(m->valBlockELL); free(m->widthVecER); free(m->rowVecER);
    free(m->biasVecER); free(m->colER); free(m->valER);
    free(m->longVecBoundary); free(m->longVecRow); free(m->longVecCol); free(m->longVecVal);
    memset(m, 0, sizeof *m);
}
