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
 *   val      double   the entries in list order (row-sorted, per-row order kept)
 *   col      uint32   the column, or 0x80000000 | k for the k-th HUB column: the (at most hubCap)
 *                     most referenced columns, whose x values every CTA keeps in shared memory -
 *                     a power-law graph sends a third of its gathers to 0.1 % of its columns
 *   grp      {seg0, mask}  per 32 entries: index of the row segment of the first entry, and bit j =
 *                     entry j starts a new row.  The segment of lane j is seg0 + popc(mask & bits
 *                     1..j): 8 bytes per 32 entries instead of 4 bytes per entry
 *   rowOfSeg[s]       the (non-empty) row of segment s
 *
 * packed TILE by TILE (ovfstream.h: the values, columns, group words and two flags of 32 x
 * tileGroups consecutive entries are one contiguous record), because the kernel moves a tile into
 * shared memory with ONE bulk copy: the first version of this format kept val / col / grp in
 * three arrays read with plain loads, and a warp then walked three DEPENDENT memory latencies per
 * tile (stream -> gather -> row of the segment): 2 278 us on R-MAT 24, behind the COO list.
 * = 12.4 bytes per entry + 4 per non-empty row (stored / algorithmic = 1.04 at R-MAT 24).  Who
 * sums what is fixed by the data: a row segment inside one warp tile is stored by one lane, the
 * pieces of a row that spans tiles go through two carry slots per tile (their rows are known here:
 * carryRow) and are added in tile order by ehyb_ovfstream_fixup - no atomics, y is bit-reproducible.
 */
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include "common.h"
#include "ovfstream.h"

void ehyb_ovfstream_free(ehyb_ovfstream *s)
{
    if (!s) return;
    free(s->tiles); free(s->rowOfSeg); free(s->hubCols); free(s->carryRow); free(s->runs);
    memset(s, 0, sizeof *s);
}

int ehyb_ovfstream_build(int64_t count, const int32_t *row, const int32_t *col, const double *val, int64_t ncols, int hubCap,
                         int tileGroups, ehyb_ovfstream *out)
{
    if (!out || count <= 0 || !row || !col || !val || ncols <= 0 || ncols > INT_MAX || (tileGroups != 4 && tileGroups != 8))
        return ehyb_fail(EHYB_ERR_ARG, "ehyb_ovfstream_build: bad argument");
    memset(out, 0, sizeof *out);
    const int TG = tileGroups;
    const int64_t E = 32 * (int64_t)TG;
    const int64_t nTiles = (count + E - 1) / E;
    const size_t tileBytes = (size_t)EHYB_OVF_TILE_BYTES(TG);
    if (nTiles > INT_MAX / 2) return ehyb_fail(EHYB_ERR_LIMIT, "overflow stream too long");
    int rc = EHYB_OK;
    int32_t *cnt = NULL, *hubIdx = NULL;
    int64_t *segAt = (int64_t *)malloc(((size_t)nTiles + 1) * sizeof(int64_t)); /* row starts before tile t */
    out->tiles = (unsigned char *)malloc((size_t)nTiles * tileBytes);
    out->carryRow = (int32_t *)malloc((size_t)nTiles * 2 * sizeof(int32_t));
    if (!segAt || !out->tiles || !out->carryRow) { rc = ehyb_fail(EHYB_ERR_NOMEM, "overflow stream: out of memory"); goto fail; }
    out->count = count; out->nTiles = nTiles; out->tileGroups = TG; out->tileBytes = (int)tileBytes;

    /* ---- hub columns: the hubCap most referenced ones (ties: smaller column first) ---- */
    int nHub = 0;
    if (hubCap > 0) {
        cnt = (int32_t *)calloc((size_t)ncols, sizeof(int32_t));
        hubIdx = (int32_t *)malloc((size_t)ncols * sizeof(int32_t));
        out->hubCols = (int32_t *)malloc((size_t)hubCap * sizeof(int32_t));
        int64_t *hist = (int64_t *)calloc(65537, sizeof(int64_t));
        if (!cnt || !hubIdx || !out->hubCols || !hist) { free(hist); rc = ehyb_fail(EHYB_ERR_NOMEM, "overflow stream: out of memory"); goto fail; }
        for (int64_t i = 0; i < count; ++i) {
            if (col[i] < 0 || col[i] >= ncols) { free(hist); rc = ehyb_fail(EHYB_ERR_ARG, "overflow list: column out of range"); goto fail; }
            cnt[col[i]] += 1;
        }
        /* threshold by a histogram of the counts (counts above 65535 share the last bin) */
        for (int64_t c = 0; c < ncols; ++c) hist[cnt[c] > 65535 ? 65536 : cnt[c]] += 1;
        int64_t acc = 0;
        int thr = 65537; /* columns with count >= thr are hubs for sure */
        for (int b = 65536; b >= 2; --b) { /* a column referenced once gains nothing from the cache */
            if (acc + hist[b] > hubCap) break;
            acc += hist[b];
            thr = b;
        }
        free(hist);
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

    /* ---- segments: row starts per tile, prefix sum ---- */
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t t = 0; t < nTiles; ++t) {
        const int64_t i0 = t * E, i1 = i0 + E < count ? i0 + E : count;
        int64_t s = 0;
        for (int64_t i = i0; i < i1; ++i) {
            if (i == 0 || row[i] != row[i - 1]) ++s;
            if (i && row[i] < row[i - 1]) bad = 1;
        }
        segAt[t + 1] = s;
    }
    if (bad) { rc = ehyb_fail(EHYB_ERR_ARG, "overflow list is not row-sorted"); goto fail; }
    segAt[0] = 0;
    for (int64_t t = 0; t < nTiles; ++t) segAt[t + 1] += segAt[t];
    const int pad = count % E != 0; /* the padding entries of the last tile form one more segment, row -1 */
    const int64_t nSeg = segAt[nTiles] + pad;
    if (nSeg > INT_MAX) { rc = ehyb_fail(EHYB_ERR_LIMIT, "overflow stream: too many rows"); goto fail; }
    out->rowOfSeg = (int32_t *)malloc((size_t)nSeg * sizeof(int32_t));
    if (!out->rowOfSeg) { rc = ehyb_fail(EHYB_ERR_NOMEM, "overflow stream: out of memory"); goto fail; }
    out->nSeg = nSeg;
    if (pad) out->rowOfSeg[nSeg - 1] = -1;

    /* ---- the tile records ---- */
    int64_t hubRefs = 0;
#pragma omp parallel for schedule(static) reduction(+ : hubRefs)
    for (int64_t t = 0; t < nTiles; ++t) {
        unsigned char *rec = out->tiles + (size_t)t * tileBytes;
        double *tv = (double *)rec;
        uint32_t *tc = (uint32_t *)(rec + 8 * E);
        uint32_t *tg = (uint32_t *)(rec + 12 * E);
        uint32_t *tf = (uint32_t *)(rec + 12 * E + 8 * TG);
        const int64_t i0 = t * E;
        int64_t seg = segAt[t] - 1; /* segment of the entry before this tile */
        for (int g = 0; g < TG; ++g) {
            uint32_t mask = 0;
            for (int j = 0; j < 32; ++j) {
                const int64_t i = i0 + 32 * g + j;
                const int k = 32 * g + j;
                if (i < count) {
                    if (i == 0 || row[i] != row[i - 1]) { ++seg; out->rowOfSeg[seg] = row[i]; mask |= 1u << j; }
                    const int32_t c = col[i];
                    if (hubIdx && hubIdx[c] >= 0) { tc[k] = EHYB_OVF_HUB_BIT | (uint32_t)hubIdx[c]; ++hubRefs; }
                    else tc[k] = (uint32_t)c;
                    tv[k] = val[i];
                } else {
                    if (i == count) { ++seg; mask |= 1u << j; } /* the padding segment (row -1): its sum is dropped */
                    tc[k] = 0; tv[k] = 0.0;
                }
                if (j == 0) tg[2 * g] = (uint32_t)seg;
            }
            tg[2 * g + 1] = mask;
        }
        const int headCont = i0 > 0 && row[i0] == row[i0 - 1];
        const int tailCont = i0 + E < count && row[i0 + E] == row[i0 + E - 1];
        tf[0] = (uint32_t)(headCont | (tailCont << 1)); tf[1] = tf[2] = tf[3] = 0;
        out->carryRow[2 * t] = headCont ? row[i0] : -1;
        out->carryRow[2 * t + 1] = tailCont ? row[i0 + E - 1] : -1;
    }
    out->hubRefs = hubRefs;
    {   /* runs of equal rows in the carry slots (a row that spans tiles), short ones first */
        int64_t nRuns = 0, nShort = 0;
        for (int pass = 0; pass < 2; ++pass) {
            int64_t ks = 0, kl = nShort;
            for (int64_t i = 0; i < 2 * nTiles;) {
                const int32_t r = out->carryRow[i];
                if (r < 0) { ++i; continue; }
                int64_t j = i + 1;
                while (j < 2 * nTiles && out->carryRow[j] == r) ++j;
                const int isShort = j - i <= EHYB_OVF_SHORT_RUN;
                if (pass == 0) { ++nRuns; nShort += isShort; }
                else {
                    int32_t *e = out->runs + 3 * (isShort ? ks++ : kl++);
                    e[0] = (int32_t)i; e[1] = (int32_t)(j - i); e[2] = r;
                }
                i = j;
            }
            if (pass == 0) {
                out->runs = (int32_t *)malloc((size_t)(nRuns ? nRuns : 1) * 3 * sizeof(int32_t));
                if (!out->runs) { rc = ehyb_fail(EHYB_ERR_NOMEM, "overflow stream: out of memory"); goto fail; }
                out->nRuns = nRuns; out->nRunsShort = nShort;
            }
        }
    }
    out->deviceBytes = nTiles * (int64_t)tileBytes + nSeg * 4 + (int64_t)nHub * 4 + out->nRuns * 12;
    free(cnt); free(hubIdx); free(segAt);
    return EHYB_OK;
fail:
    free(cnt); free(hubIdx); free(segAt);
    ehyb_ovfstream_free(out);
    return rc;
}

int ehyb_ovfstream_build_blocked(int64_t count, const int32_t *row, const int32_t *col, const double *val, int64_t ncols, int hubCap,
                                 int tileGroups, int nBlocks, int64_t blockCols, ehyb_ovfstream *out)
{
    if (!out || nBlocks < 1 || blockCols < 1 || (int64_t)nBlocks * blockCols < ncols || count <= 0 || !row || !col || !val)
        return ehyb_fail(EHYB_ERR_ARG, "ehyb_ovfstream_build_blocked: bad argument");
    memset(out, 0, sizeof *out * (size_t)nBlocks);
    if (nBlocks == 1) return ehyb_ovfstream_build(count, row, col, val, ncols, hubCap, tileGroups, out);
    int rc = EHYB_OK;
    int64_t *start = (int64_t *)calloc((size_t)nBlocks + 1, sizeof(int64_t));
    int32_t *r2 = (int32_t *)malloc((size_t)count * sizeof(int32_t)), *c2 = (int32_t *)malloc((size_t)count * sizeof(int32_t));
    double *v2 = (double *)malloc((size_t)count * sizeof(double));
    if (!start || !r2 || !c2 || !v2) { rc = ehyb_fail(EHYB_ERR_NOMEM, "overflow stream: out of memory"); goto done; }
    for (int64_t i = 0; i < count; ++i) {
        if (col[i] < 0 || col[i] >= ncols) { rc = ehyb_fail(EHYB_ERR_ARG, "overflow list: column out of range"); goto done; }
        start[col[i] / blockCols + 1] += 1;
    }
    for (int b = 0; b < nBlocks; ++b) start[b + 1] += start[b];
    {   /* stable distribution: the list order (row-sorted, per-row order) survives inside every block */
        int64_t *at = (int64_t *)malloc((size_t)nBlocks * sizeof(int64_t));
        if (!at) { rc = ehyb_fail(EHYB_ERR_NOMEM, "overflow stream: out of memory"); goto done; }
        memcpy(at, start, (size_t)nBlocks * sizeof(int64_t));
        for (int64_t i = 0; i < count; ++i) {
            const int64_t k = at[col[i] / blockCols]++;
            r2[k] = row[i]; c2[k] = col[i]; v2[k] = val[i];
        }
        free(at);
    }
    for (int b = 0; b < nBlocks && rc == EHYB_OK; ++b) {
        const int64_t cnt = start[b + 1] - start[b];
        if (cnt > 0) rc = ehyb_ovfstream_build(cnt, r2 + start[b], c2 + start[b], v2 + start[b], ncols, hubCap, tileGroups, out + b);
    }
    if (rc) for (int b = 0; b < nBlocks; ++b) ehyb_ovfstream_free(out + b);
done:
    free(start); free(r2); free(c2); free(v2);
    return rc;
}
