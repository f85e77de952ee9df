/*
 * ovfstream.c -- device format of a LARGE overflow list: a row-sorted CSR-like stream.
 *
 * The overflow list of a layout is a COO (row, col, val), 16 bytes per entry, reduced on the device
 * with one atomicAdd per row segment (ehyb_overflow_kernel).  That is fine for what the list was
 * made for - the few entries the slices cannot hold, halo entries - but for matrices the explicit
 * cache cannot help (power-law graphs: BASELINE.json config 4, R-MAT scale 24, keeps every entry
 * there) the list IS the matrix: 1.23 x the algorithmic bytes, L1TEX-bound on 32 uncoalesced sectors
 * per warp gather, behind cuSPARSE CSR, and y not bit-reproducible.  This file
 * turns the list into what the reference's remainder intent asks for (kernel.cu:43-67 long rows,
 * :80-108 ER: one owner per row, sums by warp shuffles) in a form a stream kernel can eat:
 *
 *   val[i]   double   the entries in list order (row-sorted, per-row order kept)
 *   col[i]   uint32   the column, or 0x80000000 | k for the k-th HUB column: the (at most hubCap)
 *                     most referenced columns, whose x values every CTA keeps in shared memory -
 *                     a power-law graph sends a third of its gathers to 0.1 % of its columns
 *   grp[g]   {seg0, mask}  per 32 entries: index of the row segment of entry 32 g, and bit j = entry
 *                     32 g + j starts a new row.  The segment of lane j is seg0 + popc(mask & bits
 *                     1..j): 8 bytes per 32 entries instead of 4 bytes per entry
 *   rowOfSeg[s]       the (non-empty) row of segment s
 *
 * = 12.25 bytes per entry + 4 per non-empty row (stored / algorithmic = 1.03 at R-MAT 24).  Who
 * sums what is fixed by the data: a row segment inside one warp tile is stored by one lane, the
 * pieces of a row that spans tiles go through two carry slots per tile and are added in tile order
 * by ehyb_ovfstream_fixup - no atomics, y is bit-reproducible.
 */
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include "common.h"
#include "ovfstream.h"

void ehyb_ovfstream_free(ehyb_ovfstream *s)
{
    if (!s) return;
    free(s->col); free(s->grp); free(s->rowOfSeg); free(s->hubCols);
    memset(s, 0, sizeof *s);
}

int ehyb_ovfstream_build(int64_t count, const int32_t *row, const int32_t *col, int64_t ncols, int hubCap, ehyb_ovfstream *out)
{
    if (!out || count <= 0 || !row || !col || ncols <= 0 || ncols > INT_MAX) return ehyb_fail(EHYB_ERR_ARG, "ehyb_ovfstream_build: bad argument");
    memset(out, 0, sizeof *out);
    const int64_t nGroups = (count + 31) / 32;
    if (nGroups > INT_MAX) return ehyb_fail(EHYB_ERR_LIMIT, "overflow stream too long");
    int rc = EHYB_OK;
    int32_t *cnt = (int32_t *)calloc((size_t)ncols, sizeof(int32_t));
    int32_t *hubIdx = NULL;
    out->col = (uint32_t *)malloc((size_t)count * sizeof(uint32_t));
    out->grp = (uint32_t *)malloc((size_t)nGroups * 2 * sizeof(uint32_t));
    if (!cnt || !out->col || !out->grp) { rc = ehyb_fail(EHYB_ERR_NOMEM, "overflow stream: out of memory"); goto fail; }
    out->count = count; out->nGroups = nGroups;

    /* ---- hub columns: the hubCap most referenced ones (ties: smaller column first) ---- */
    int nHub = 0;
    if (hubCap > 0) {
        for (int64_t i = 0; i < count; ++i) cnt[col[i]] += 1;
        /* threshold by a histogram of the counts (counts above 65535 share the last bin) */
        int64_t *hist = (int64_t *)calloc(65537, sizeof(int64_t));
        if (!hist) { rc = ehyb_fail(EHYB_ERR_NOMEM, "overflow stream: out of memory"); goto fail; }
        for (int64_t c = 0; c < ncols; ++c) hist[cnt[c] > 65535 ? 65536 : cnt[c]] += 1;
        int64_t acc = 0;
        int thr = 65537; /* columns with count >= thr are hubs for sure */
        for (int b = 65536; b >= 2; --b) { /* a column referenced once gains nothing from the cache */
            if (acc + hist[b] > hubCap) break;
            acc += hist[b];
            thr = b;
        }
        free(hist);
        out->hubCols = (int32_t *)malloc((size_t)(hubCap > 0 ? hubCap : 1) * sizeof(int32_t));
        hubIdx = (int32_t *)malloc((size_t)ncols * sizeof(int32_t));
        if (!out->hubCols || !hubIdx) { rc = ehyb_fail(EHYB_ERR_NOMEM, "overflow stream: out of memory"); goto fail; }
        /* everything >= thr, then fill the remaining places with columns of count thr - 1 in column order */
        for (int64_t c = 0; c < ncols; ++c) {
            hubIdx[c] = -1;
            if (thr <= 65536 && cnt[c] >= thr) { hubIdx[c] = nHub; out->hubCols[nHub++] = (int32_t)c; }
        }
        if (thr > 2 && thr <= 65537) {
            const int fill = thr - 1 > 65535 ? 0 : thr - 1; /* the shared last bin is not split */
            for (int64_t c = 0; c < ncols && nHub < hubCap && fill >= 2; ++c)
                if (cnt[c] == fill) { hubIdx[c] = nHub; out->hubCols[nHub++] = (int32_t)c; }
        }
    }
    out->nHub = nHub;

    /* ---- columns, group descriptors, segments ---- */
    int64_t nSeg = 0, hubRefs = 0;
    for (int64_t i = 0; i < count; ++i)
        if (i == 0 || row[i] != row[i - 1]) {
            if (i && row[i] < row[i - 1]) { rc = ehyb_fail(EHYB_ERR_ARG, "overflow list is not row-sorted"); goto fail; }
            ++nSeg;
        }
    if (nSeg > INT_MAX) { rc = ehyb_fail(EHYB_ERR_LIMIT, "overflow stream: too many rows"); goto fail; }
    out->rowOfSeg = (int32_t *)malloc((size_t)nSeg * sizeof(int32_t));
    if (!out->rowOfSeg) { rc = ehyb_fail(EHYB_ERR_NOMEM, "overflow stream: out of memory"); goto fail; }
    out->nSeg = nSeg;
    int64_t seg = -1;
    for (int64_t g = 0; g < nGroups; ++g) {
        uint32_t mask = 0;
        const int64_t i0 = g * 32, i1 = i0 + 32 < count ? i0 + 32 : count;
        for (int64_t i = i0; i < i1; ++i) {
            const int start = i == 0 || row[i] != row[i - 1];
            if (start) { ++seg; out->rowOfSeg[seg] = row[i]; mask |= 1u << (i - i0); }
            if (i == i0) out->grp[2 * g] = (uint32_t)seg;
            const int32_t c = col[i];
            if (hubIdx && hubIdx[c] >= 0) { out->col[i] = 0x80000000u | (uint32_t)hubIdx[c]; ++hubRefs; }
            else out->col[i] = (uint32_t)c;
        }
        out->grp[2 * g + 1] = mask;
    }
    out->hubRefs = hubRefs;
    out->deviceBytes = count * 12 + nGroups * 8 + nSeg * 4 + (int64_t)nHub * 4;
    free(cnt); free(hubIdx);
    return EHYB_OK;
fail:
    free(cnt); free(hubIdx);
    ehyb_ovfstream_free(out);
    return rc;
}
