/*
 * layout.c -- the Blackwell-tuned EHYB layout (host-side build, view, de-interleave).
 *
 * Logical content is the reference's (convert.c:247-267): an entry of permuted row r in
 * partition p is an ELL entry iff partStart <= J < partStart+W, stored with the 16-bit
 * window-local column J-partStart; everything else is remainder, stored with the 32-bit
 * permuted column.  Per-row entry order is kept in both parts.  What changes is the tiling:
 *
 *   slice      64 consecutive rows of one partition (never straddles a partition).  Lane l of
 *              the warp that processes the slice owns rows l and l+32 ("halves" h=0,1).
 *   slice blob [ELL values ][ELL columns ][remainder values ][remainder columns]
 *                w*512 B     ceil4(w)*128 B   wr*512 B          ceil4(wr)*128 B
 *              values         double  [k][lane][h]         one 128-bit access per lane per k
 *              columns        uint16  [k/4][lane][h][k%4]  one 128-bit access per lane per 4 k
 *              ELL columns index the x window of the partition (J - partStart); remainder
 *              columns index the partition's REMAINDER CACHE: the ascending list of the (at
 *              most cache_cap) most referenced columns outside the window, whose x values the
 *              kernel gathers into shared memory once per CTA - the explicit cache extended
 *              to the remainder, so that the steady-state loop never gathers from global
 *              memory (the L1 tag stage serialises 32 uncoalesced sectors per warp load).
 *              Every region is a multiple of 256 bytes, so slices start 256-byte aligned and
 *              a slice is addressed by a 32-bit offset in 256-byte units.
 *   w          max in-window count over the slice's rows (the reference's rule, per 64 rows)
 *   wr         in-slice remainder width: the ceil(er_fill*64)-th largest count of cached
 *              remainder entries in the slice (er_fill=0: the largest).  Remainder entries
 *              beyond wr or with an uncached column, and whole "long" rows (more than
 *              long_row_threshold in-window entries at a partition head, the reference's
 *              longVec rule), go to the overflow list.
 *   overflow   COO (row, col, val), sorted by row, entry order kept; reduced on the device
 *              by a segmented warp reduction and added to y.
 *
 * Unlike the reference layout (W/32 slices per partition, rows beyond the window in a
 * separate global list) every row of a partition belongs to a slice here; a row beyond the
 * window has no ELL entries, all of its entries are remainder - the reference's
 * classification (convert.c:128-134, :285-306).
 */
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.h"
#include "convert.h"
#include "kernel.h"

#define SR EHYB_SLICE_ROWS
#define EHYB_BALANCE_WARPS 24 /* warps of a staged-kernel CTA (kMaxStageWarps) */

struct ehyb_layout {
    ehyb_layout_view v;
    ehyb_part_desc *parts;
    ehyb_slice_desc *slices;
    unsigned char *blob;
    int32_t *ovfRow, *ovfCol;
    double *ovfVal;
    /* host-only bookkeeping for the de-interleave */
    int32_t *rowEll;   /* [n] ELL entries of the row */
    int32_t *rowRemIn; /* [n] remainder entries kept in the slice */
    int32_t *rowCached; /* [n] remainder entries whose column is in the partition's cache */
    int32_t *cacheCols; /* concatenated per-partition cache lists (ascending permuted columns) */
    int64_t *ovfPtr;   /* [n+1] overflow entries of the row */
    int blobBorrowed;  /* the blob belongs to a streamed builder (a chunk written in place) */
};

int ehyb_convert_reference_layout(const matrixCOO *in, matrixEHYB *out, int *sizeBlockELL, int *sizeER, int quiet);

void ehyb_layout_free(ehyb_layout *L)
{
    if (!L) return;
    free(L->parts); free(L->slices);
    if (!L->blobBorrowed) free(L->blob);
    free(L->ovfRow); free(L->ovfCol); free(L->ovfVal);
    free(L->rowEll); free(L->rowRemIn); free(L->rowCached); free(L->cacheCols); free(L->ovfPtr);
    free(L);
}

int ehyb_layout_get(const ehyb_layout *L, ehyb_layout_view *view)
{
    if (!L || !view) return ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_get: NULL argument");
    *view = L->v;
    return EHYB_OK;
}

/* byte offsets of the four regions inside a slice */
static inline int64_t reg_ell_col(int w) { return (int64_t)w * 512; }
static inline int64_t reg_rem_val(int w) { return (int64_t)w * 512 + (int64_t)((w + 3) / 4) * 512; }
static inline int64_t reg_rem_col(int w, int wr) { return reg_rem_val(w) + (int64_t)wr * 512; }
static inline int64_t slice_bytes(int w, int wr) { return reg_rem_col(w, wr) + (int64_t)((wr + 3) / 4) * 512; }

static int cmp_int_desc(const void *a, const void *b) { return *(const int *)b - *(const int *)a; }

typedef struct { int32_t col; int32_t cnt; } col_count;

static int cmp_i32(const void *a, const void *b)
{
    const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return x < y ? -1 : x > y;
}

static int cmp_col_count(const void *a, const void *b) /* more references first, then column */
{
    const col_count *x = (const col_count *)a, *y = (const col_count *)b;
    if (x->cnt != y->cnt) return x->cnt > y->cnt ? -1 : 1;
    return x->col < y->col ? -1 : x->col > y->col;
}

/* position of c in the ascending list, or -1 */
static inline int cache_find(const int32_t *list, int count, int32_t c)
{
    int lo = 0, hi = count - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (list[mid] == c) return mid;
        if (list[mid] < c) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

/* where a chunk's slice data goes when the caller (the streamed builder) owns one big blob:
 * zero-filled memory of `cap` bytes; used if the chunk fits, else the chunk allocates its own */
typedef struct { unsigned char *dst; int64_t cap; int used; } blob_target;

static int layout_build_impl(int64_t nGlobal64, int64_t rowBase64, int64_t n64, const int64_t *rowPtr, const int32_t *col, const double *val,
                             int nParts, const int32_t *pb, const ehyb_layout_opts *opts, int allOverflow, blob_target *bt, ehyb_layout **out);

/*
 * Format decision for matrices the explicit cache cannot help (power-law graphs: R-MAT scale 24
 * keeps 1.9 % of its entries in the window or the remainder cache): when less than
 * min_coverage of the entries would live in slices, the slices are pure overhead (2 072
 * partitions x window + cache staging for nothing: 275 us of a 1 867 us product), so ALL entries
 * go to the row-sorted COO list and the device session replaces the main kernel by a memset.
 */
int ehyb_layout_build_csr(int64_t n64, const int64_t *rowPtr, const int32_t *col, const double *val, int nParts,
                          const int32_t *pb, const ehyb_layout_opts *opts, ehyb_layout **out)
{
    if (!opts || !out) return ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_build: bad argument");
    int rc = layout_build_impl(n64, 0, n64, rowPtr, col, val, nParts, pb, opts, 0, NULL, out);
    if (rc) return rc;
    const double minCov = opts->min_coverage > 0 ? opts->min_coverage : (opts->min_coverage < 0 ? 0.0 : EHYB_DEFAULT_MIN_COVERAGE);
    const ehyb_layout_view *v = &(*out)->v;
    /* (not for distributed blocks built for the peer-memory exchange: there the main kernel
     * carries the halo push, and the stencil/FEM blocks it is meant for are far above the limit) */
    if (v->nnz > 0 && (double)(v->nnzEll + v->nnzRemInSlice) < minCov * (double)v->nnz && v->ncols == v->n) {
        ehyb_layout_free(*out);
        *out = NULL;
        rc = layout_build_impl(n64, 0, n64, rowPtr, col, val, nParts, pb, opts, 1, NULL, out);
    }
    return rc;
}

/* Builds the layout of the rows [rowBase, rowBase + nRows) of a matrix with nGlobal rows: either
 * the whole matrix (rowBase 0, nRows == nGlobal) or one chunk of consecutive partitions of it
 * (streamed build, ehyb_layout_builder_add).  rowPtr[nRows + 1] is relative to the chunk
 * (rowPtr[0] == 0), pb[nParts + 1] and the columns are absolute.  A chunk's layout numbers its
 * slices, cache lists and blob offsets from zero (the builder shifts them when it appends). */
static int layout_build_impl(int64_t nGlobal64, int64_t rowBase64, int64_t n64, const int64_t *rowPtr, const int32_t *col, const double *val,
                             int nParts, const int32_t *pb, const ehyb_layout_opts *opts, int allOverflow, blob_target *bt, ehyb_layout **out)
{
    if (!rowPtr || !col || !val || !pb || !opts || !out || n64 <= 0 || nGlobal64 > INT_MAX || rowBase64 < 0 || rowBase64 + n64 > nGlobal64 || nParts <= 0)
        return ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_build: bad argument");
    const int n = (int)nGlobal64, rb = (int)rowBase64, nRows = (int)n64, P = nParts, W = opts->W;
    const int isChunk = nRows != n;
    const int64_t ncols = opts->ncols > 0 ? opts->ncols : n;
    const int longThr = opts->long_row_threshold > 0 ? opts->long_row_threshold : EHYB_REF_LONG_ROW;
    const double fill = opts->er_fill < 0 ? 0.0 : opts->er_fill;
    int need = (int)ceil(fill * SR);
    if (need < 1) need = 1;
    if (need > SR) need = SR;
    /* remainder cache capacity: explicit, none (< 0), or what a B200 CTA can hold next to the
     * window while leaving ~100 KB of staging slots for the matrix stream */
    int cacheCap = opts->cache_cap > 0 ? opts->cache_cap : 0;
    if (opts->cache_cap == 0) {
        const long room = 232448L - 512 - ((long)W + 2) * 8 - 100 * 1024;
        cacheCap = room > (long)EHYB_DEFAULT_CACHE_CAP * 8 ? (int)(room / 8) : EHYB_DEFAULT_CACHE_CAP;
    }
    if (W <= 0 || W > 65536) return ehyb_fail(EHYB_ERR_LIMIT, "window %d outside (0, 65536]: column indices are 16-bit", W);
    if (cacheCap > 65536) cacheCap = 65536;
    if (ncols < n || ncols > INT_MAX) return ehyb_fail(EHYB_ERR_ARG, "ncols %lld must be in [n, 2^31)", (long long)ncols);
    if (pb[0] != rb || pb[P] != rb + nRows) return ehyb_fail(EHYB_ERR_ARG, "partBoundary must run from %d to %d", rb, rb + nRows);
    for (int p = 0; p < P; ++p)
        if (pb[p] > pb[p + 1]) return ehyb_fail(EHYB_ERR_ARG, "partBoundary not monotone at %d", p);

    ehyb_layout *L = (ehyb_layout *)calloc(1, sizeof *L);
    if (!L) return ehyb_fail(EHYB_ERR_NOMEM, "layout: out of memory");
    int rc = EHYB_OK;
    int64_t *sliceOff = NULL;
    int32_t **cacheList = (int32_t **)calloc((size_t)P, sizeof(int32_t *)); /* per partition, ascending */
    int32_t *wrCap = (int32_t *)malloc((size_t)P * sizeof(int32_t));          /* per partition: widest in-slice remainder */
    L->parts = (ehyb_part_desc *)calloc((size_t)P, sizeof(ehyb_part_desc));
    L->rowEll = (int32_t *)calloc((size_t)nRows, sizeof(int32_t));
    L->rowRemIn = (int32_t *)calloc((size_t)nRows, sizeof(int32_t));
    L->rowCached = (int32_t *)calloc((size_t)nRows, sizeof(int32_t));
    L->ovfPtr = (int64_t *)calloc((size_t)nRows + 1, sizeof(int64_t));
    if (!cacheList || !wrCap || !L->parts || !L->rowEll || !L->rowRemIn || !L->rowCached || !L->ovfPtr) { rc = ehyb_fail(EHYB_ERR_NOMEM, "layout: out of memory"); goto fail; }
    /* distributed blocks: an entry whose column is in the halo [n, ncols) arrives with the
     * exchange, after the main kernel has started: it is never cached, always overflow */
    const int haloOvf = opts->halo_in_overflow && ncols > n;

    int nSlices = 0;
    for (int p = 0; p < P; ++p) {
        L->parts[p].rowStart = pb[p];
        L->parts[p].rowEnd = pb[p + 1];
        L->parts[p].sliceStart = nSlices;
        nSlices += (pb[p + 1] - pb[p] + SR - 1) / SR;
        L->parts[p].sliceEnd = nSlices;
    }
    L->slices = (ehyb_slice_desc *)calloc((size_t)(nSlices ? nSlices : 1), sizeof(ehyb_slice_desc));
    sliceOff = (int64_t *)calloc((size_t)nSlices + 1, sizeof(int64_t));
    if (!L->slices || !sliceOff) { rc = ehyb_fail(EHYB_ERR_NOMEM, "layout: out of memory"); goto fail; }

    /* pass 1a, per partition: classify the entries of every row (ELL count; -1 marks a long
     * row), then choose the remainder cache: the cacheCap most referenced columns outside the
     * window, kept in ascending order (position = 16-bit index stored in the slices) */
    int64_t nnzEll = 0, nnzRemIn = 0, nnzOvf = 0, padEll = 0, padRem = 0, nLong = 0;
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nLong) reduction(| : bad)
    for (int p = 0; p < P; ++p) {
        const int ps = pb[p], pe = pb[p + 1];
        const int64_t winEnd = (int64_t)ps + W < n ? (int64_t)ps + W : n;
        int firstReg = ps, scanning = 1;
        int64_t nExt = 0;
        /* Long-row limit of this partition.  The reference's rule is a constant (512 in-window
         * entries, kernel.h:26); a slice is walked by ONE warp, chunk after chunk, so a slice much
         * wider than a warp's fair share of the partition serialises the CTA (R-MAT: 106 us for
         * 2 M entries).  The limit is therefore also bounded by twice the columns a warp would
         * own if the partition's in-window entries were spread evenly over EHYB_BALANCE_WARPS
         * warps (never below 32).  Regular matrices are unaffected: a 27-point stencil partition
         * holds ~110 columns per warp and no row has more than 27 entries. */
        int partThr = longThr;
        int64_t fairCols = -1;
        wrCap[p] = 65535;
        if (opts->long_row_threshold <= 0) {
            int64_t inWin = 0;
            for (int r = ps; r < pe; ++r) {
                if (r - ps >= W) break;
                for (int64_t e = rowPtr[r - rb]; e < rowPtr[r + 1 - rb]; ++e) inWin += (col[e] >= ps && col[e] < winEnd);
            }
            const int64_t fair = inWin / SR / EHYB_BALANCE_WARPS; /* columns per warp, evenly spread */
            int64_t lim = 2 * fair;
            if (lim < 32) lim = 32;
            if (lim < partThr) partThr = (int)lim;
            fairCols = fair;
        }
        for (int r = ps; r < pe; ++r) {
            int ell = 0;
            const int inWindowRow = r - ps < W; /* rows beyond the window are remainder as a whole (convert.c:128-134) */
            for (int64_t e = rowPtr[r - rb]; e < rowPtr[r + 1 - rb]; ++e) {
                const int c = col[e];
                if (c < 0 || c >= ncols) bad = 1;
                ell += (inWindowRow && c >= ps && c < winEnd);
            }
            if (allOverflow || (scanning && ell > partThr)) { /* long rows sit at the head of the partition */
                firstReg = r + 1;
                L->rowEll[r - rb] = -1;
                continue;
            }
            scanning = 0;
            L->rowEll[r - rb] = ell;
            nExt += rowPtr[r + 1 - rb] - rowPtr[r - rb] - ell;
        }
        nLong += firstReg - ps;
        if (cacheCap == 0 || nExt == 0 || bad) continue;
        int32_t *ext = (int32_t *)malloc((size_t)nExt * sizeof(int32_t));
        if (!ext) { bad = 2; continue; }
        int64_t m = 0;
        for (int r = firstReg; r < pe; ++r) {
            const int inWindowRow = r - ps < W;
            for (int64_t e = rowPtr[r - rb]; e < rowPtr[r + 1 - rb]; ++e) {
                const int c = col[e];
                if (inWindowRow && c >= ps && c < winEnd) continue;
                if (haloOvf && c >= n) continue;
                ext[m++] = c;
            }
        }
        qsort(ext, (size_t)m, sizeof(int32_t), cmp_i32);
        int64_t u = 0;
        for (int64_t i = 0; i < m; ++i)
            if (i == 0 || ext[i] != ext[i - 1]) ++u;
        int32_t *chosen;
        int nChosen;
        if (u <= cacheCap) { /* everything fits: the unique list itself */
            chosen = ext;
            nChosen = 0;
            for (int64_t i = 0; i < m; ++i)
                if (i == 0 || ext[i] != ext[i - 1]) chosen[nChosen++] = ext[i];
        } else {
            col_count *cc = (col_count *)malloc((size_t)u * sizeof(col_count));
            if (!cc) { free(ext); bad = 2; continue; }
            int64_t k = -1;
            for (int64_t i = 0; i < m; ++i) {
                if (i == 0 || ext[i] != ext[i - 1]) { ++k; cc[k].col = ext[i]; cc[k].cnt = 0; }
                cc[k].cnt += 1;
            }
            qsort(cc, (size_t)u, sizeof(col_count), cmp_col_count);
            chosen = ext;
            nChosen = cacheCap;
            for (int i = 0; i < nChosen; ++i) chosen[i] = cc[i].col;
            qsort(chosen, (size_t)nChosen, sizeof(int32_t), cmp_i32);
            free(cc);
        }
        cacheList[p] = chosen;
        L->parts[p].cacheCount = nChosen;
        for (int r = firstReg; r < pe; ++r) {
            const int inWindowRow = r - ps < W;
            /* the in-slice part of a row's remainder is a PREFIX of its remainder sequence
             * (so that in-slice + overflow, concatenated, is the original order): count the
             * leading remainder entries whose column is cached */
            int cached = 0;
            for (int64_t e = rowPtr[r - rb]; e < rowPtr[r + 1 - rb]; ++e) {
                const int c = col[e];
                if (inWindowRow && c >= ps && c < winEnd) continue;
                if ((haloOvf && c >= n) || cache_find(chosen, nChosen, c) < 0) break;
                cached += 1;
            }
            L->rowCached[r - rb] = cached;
        }
        /* the same balance rule for the in-slice remainder width: at most twice a warp's fair
         * share of all the partition's columns (ELL + cached remainder), never below 32 */
        if (fairCols >= 0) {
            int64_t sumCached = 0;
            for (int r = firstReg; r < pe; ++r) sumCached += L->rowCached[r - rb];
            int64_t lim = 2 * (fairCols + sumCached / SR / EHYB_BALANCE_WARPS);
            if (lim < 32) lim = 32;
            if (lim < wrCap[p]) wrCap[p] = (int32_t)lim;
        }
    }
    if (bad) { rc = bad == 2 ? ehyb_fail(EHYB_ERR_NOMEM, "layout: out of memory") : ehyb_fail(EHYB_ERR_ARG, "layout: a column index is outside [0, ncols)"); goto fail; }
    int64_t cacheTotal = 0;
    int cacheMax = 0;
    for (int p = 0; p < P; ++p) {
        L->parts[p].cacheStart = (int32_t)cacheTotal;
        cacheTotal += L->parts[p].cacheCount;
        if (L->parts[p].cacheCount > cacheMax) cacheMax = L->parts[p].cacheCount;
    }
    if (cacheTotal > INT_MAX) { rc = ehyb_fail(EHYB_ERR_LIMIT, "remainder cache lists exceed 2^31 entries"); goto fail; }
    L->cacheCols = (int32_t *)malloc((size_t)(cacheTotal ? cacheTotal : 1) * sizeof(int32_t));
    if (!L->cacheCols) { rc = ehyb_fail(EHYB_ERR_NOMEM, "layout: out of memory"); goto fail; }
    for (int p = 0; p < P; ++p)
        if (L->parts[p].cacheCount) memcpy(L->cacheCols + L->parts[p].cacheStart, cacheList[p], (size_t)L->parts[p].cacheCount * sizeof(int32_t));

    /* er_fill < 0: choose between "every cached remainder entry in its slice" (need = 1) and "a
     * remainder column only while half of the lanes use it" (need = 32) by stored bytes:
     * 10 B per in-slice slot (padding included), 16 B per overflow entry plus the cost of
     * launching the overflow kernel at all (~5 us of streaming, 30 MB). */
    if (opts->er_fill < 0) {
        double cost[2] = {0.0, 0.0};
        int64_t ovfB = 0, ovfAlways = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : cost[:2], ovfB, ovfAlways)
        for (int p = 0; p < P; ++p) {
            const int ps = pb[p], pe = pb[p + 1];
            for (int s = L->parts[p].sliceStart; s < L->parts[p].sliceEnd; ++s) {
                const int r0 = ps + (s - L->parts[p].sliceStart) * SR;
                const int r1 = r0 + SR < pe ? r0 + SR : pe;
                int m = 0, rem[SR];
                for (int r = r0; r < r1; ++r) {
                    const int64_t len = rowPtr[r + 1 - rb] - rowPtr[r - rb];
                    if (L->rowEll[r - rb] < 0) { ovfAlways += len; continue; }
                    rem[m++] = L->rowCached[r - rb];
                    ovfAlways += len - L->rowEll[r - rb] - L->rowCached[r - rb];
                }
                qsort(rem, (size_t)m, sizeof(int), cmp_int_desc);
                int wrA = m >= 1 ? rem[0] : 0, wrB = m >= SR / 2 ? rem[SR / 2 - 1] : 0;
                if (wrA > wrCap[p]) wrA = wrCap[p];
                if (wrB > wrCap[p]) wrB = wrCap[p];
                cost[0] += 640.0 * wrA;
                cost[1] += 640.0 * wrB;
                for (int i = 0; i < m; ++i) {
                    if (rem[i] > wrA) cost[0] += 16.0 * (rem[i] - wrA);
                    if (rem[i] > wrB) { cost[1] += 16.0 * (rem[i] - wrB); ovfB += rem[i] - wrB; }
                }
            }
        }
        if (ovfB > 0 && ovfAlways == 0) cost[1] += 30e6; /* the overflow launch would exist only because of this choice */
        need = cost[0] <= cost[1] ? 1 : SR / 2;
    }

    /* pass 1b: slice widths, in-slice / overflow split */
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nnzEll, nnzRemIn, nnzOvf, padEll, padRem) reduction(| : bad)
    for (int p = 0; p < P; ++p) {
        const int ps = pb[p], pe = pb[p + 1];
        for (int s = L->parts[p].sliceStart; s < L->parts[p].sliceEnd; ++s) {
            const int r0 = ps + (s - L->parts[p].sliceStart) * SR;
            const int r1 = r0 + SR < pe ? r0 + SR : pe;
            int w = 0, m = 0, rem[SR];
            for (int r = r0; r < r1; ++r) {
                if (L->rowEll[r - rb] < 0) continue;
                if (L->rowEll[r - rb] > w) w = L->rowEll[r - rb];
                rem[m++] = L->rowCached[r - rb] > 65535 ? 65535 : L->rowCached[r - rb];
            }
            qsort(rem, (size_t)m, sizeof(int), cmp_int_desc);
            int wr = m >= need ? rem[need - 1] : 0;
            if (wr > wrCap[p]) wr = wrCap[p];
            if (w > 65535) bad = 1;
            L->slices[s].w = (uint16_t)w;
            L->slices[s].wr = (uint16_t)wr;
            sliceOff[s + 1] = slice_bytes(w, wr);
            int64_t sumE = 0, sumR = 0;
            for (int r = r0; r < r1; ++r) {
                const int64_t len = rowPtr[r + 1 - rb] - rowPtr[r - rb];
                if (L->rowEll[r - rb] < 0) { /* long row: everything overflows */
                    L->ovfPtr[r + 1 - rb] = len;
                    nnzOvf += len;
                    continue;
                }
                const int in = L->rowCached[r - rb] < wr ? L->rowCached[r - rb] : wr;
                L->rowRemIn[r - rb] = in;
                L->ovfPtr[r + 1 - rb] = len - L->rowEll[r - rb] - in;
                sumE += L->rowEll[r - rb];
                sumR += in;
                nnzOvf += len - L->rowEll[r - rb] - in;
            }
            nnzEll += sumE;
            nnzRemIn += sumR;
            padEll += (int64_t)SR * w - sumE;
            padRem += (int64_t)SR * wr - sumR;
        }
    }
    if (bad) { rc = ehyb_fail(EHYB_ERR_LIMIT, "layout: a slice is wider than 65535"); goto fail; }

    for (int s = 0; s < nSlices; ++s) {
        if (sliceOff[s] / 256 > (int64_t)UINT32_MAX) { rc = ehyb_fail(EHYB_ERR_LIMIT, "layout larger than 1 TiB"); goto fail; }
        L->slices[s].off256 = (uint32_t)(sliceOff[s] / 256);
        sliceOff[s + 1] += sliceOff[s];
    }
    const int64_t blobBytes = sliceOff[nSlices];
    for (int r = 0; r < nRows; ++r) L->ovfPtr[r + 1] += L->ovfPtr[r];
    const int64_t nOvf = L->ovfPtr[nRows];
    if (nOvf > INT_MAX) { rc = ehyb_fail(EHYB_ERR_LIMIT, "overflow list exceeds 2^31 entries"); goto fail; }

    if (bt && bt->dst && blobBytes <= bt->cap) { /* straight into the caller's blob (already zero) */
        L->blob = bt->dst;
        L->blobBorrowed = 1;
        bt->used = 1;
    } else {
        L->blob = (unsigned char *)calloc((size_t)(blobBytes ? blobBytes : 256), 1);
    }
    L->ovfRow = (int32_t *)malloc((size_t)(nOvf ? nOvf : 1) * sizeof(int32_t));
    L->ovfCol = (int32_t *)malloc((size_t)(nOvf ? nOvf : 1) * sizeof(int32_t));
    L->ovfVal = (double *)malloc((size_t)(nOvf ? nOvf : 1) * sizeof(double));
    if (!L->blob || !L->ovfRow || !L->ovfCol || !L->ovfVal) { rc = ehyb_fail(EHYB_ERR_NOMEM, "layout: out of memory (%lld bytes)", (long long)blobBytes); goto fail; }

    /* pass 2: fill.  Padding stays zero (value 0.0, index 0: a valid window / cache position). */
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < P; ++p) {
        const int ps = pb[p], pe = pb[p + 1];
        const int64_t winEnd = (int64_t)ps + W < n ? (int64_t)ps + W : n;
        const int32_t *cl = cacheList[p];
        const int cn = L->parts[p].cacheCount;
        for (int s = L->parts[p].sliceStart; s < L->parts[p].sliceEnd; ++s) {
            const int r0 = ps + (s - L->parts[p].sliceStart) * SR;
            const int r1 = r0 + SR < pe ? r0 + SR : pe;
            const int w = L->slices[s].w, wr = L->slices[s].wr;
            unsigned char *base = L->blob + sliceOff[s];
            double *ev = (double *)base;
            uint16_t *ec = (uint16_t *)(base + reg_ell_col(w));
            double *rv = (double *)(base + reg_rem_val(w));
            uint16_t *rcol = (uint16_t *)(base + reg_rem_col(w, wr));
            for (int r = r0; r < r1; ++r) {
                const int t = r - r0, lane = t % 32, h = t / 32;
                const int isLong = L->rowEll[r - rb] < 0;
                const int inWindowRow = !isLong && r - ps < W;
                int kE = 0, kR = 0;
                int64_t o = L->ovfPtr[r - rb];
                for (int64_t e = rowPtr[r - rb]; e < rowPtr[r + 1 - rb]; ++e) {
                    const int c = col[e];
                    int ci = -1;
                    if (inWindowRow && c >= ps && c < winEnd) {
                        ev[((int64_t)kE * 32 + lane) * 2 + h] = val[e];
                        ec[(((int64_t)(kE / 4) * 32 + lane) * 2 + h) * 4 + kE % 4] = (uint16_t)(c - ps);
                        ++kE;
                    } else if (!isLong && kR < L->rowRemIn[r - rb] && (ci = cache_find(cl, cn, c)) >= 0) {
                        rv[((int64_t)kR * 32 + lane) * 2 + h] = val[e];
                        rcol[(((int64_t)(kR / 4) * 32 + lane) * 2 + h) * 4 + kR % 4] = (uint16_t)ci;
                        ++kR;
                    } else {
                        L->ovfRow[o] = r;
                        L->ovfCol[o] = c;
                        L->ovfVal[o] = val[e];
                        ++o;
                    }
                }
            }
        }
    }
    for (int r = 0; r < nRows; ++r)
        if (L->rowEll[r] < 0) L->rowEll[r] = 0; /* long rows own no ELL entries */

    ehyb_layout_view *v = &L->v;
    v->n = nRows; v->ncols = ncols; v->nnz = rowPtr[nRows];
    v->nParts = P; v->W = W; v->ctasPerPart = opts->ctasPerPart > 0 ? opts->ctasPerPart : 1; v->nSlices = nSlices;
    v->parts = L->parts; v->slices = L->slices; v->blob = L->blob; v->blobBytes = blobBytes;
    v->nOverflow = nOvf; v->ovfRow = L->ovfRow; v->ovfCol = L->ovfCol; v->ovfVal = L->ovfVal;
    v->cacheCols = L->cacheCols; v->cacheTotal = cacheTotal; v->cacheMax = cacheMax;
    v->haloInOverflow = haloOvf;
    v->nnzEll = nnzEll; v->nnzRemInSlice = nnzRemIn; v->nnzOverflow = nnzOvf;
    v->padEll = padEll; v->padRem = padRem; v->nLongRows = nLong;
    v->algBytes = 8 * v->nnz + 2 * nnzEll + 4 * (v->nnz - nnzEll) + (isChunk ? 0 : 8 * ncols + 8 * (int64_t)n); /* x and y once per matrix */
    v->formatBytes = blobBytes + (int64_t)nSlices * (int64_t)sizeof(ehyb_slice_desc) +
                     (int64_t)P * (int64_t)sizeof(ehyb_part_desc) + nOvf * 16 + cacheTotal * 4;
    if (nnzEll + nnzRemIn + nnzOvf != v->nnz) { rc = ehyb_fail(EHYB_ERR_ARG, "layout: entry count mismatch"); goto fail; }
    for (int p = 0; p < P; ++p) free(cacheList[p]);
    free(cacheList);
    free(wrCap);
    free(sliceOff);
    *out = L;
    return EHYB_OK;

fail:
    if (cacheList) for (int p = 0; p < P; ++p) free(cacheList[p]);
    free(cacheList);
    free(wrCap);
    free(sliceOff);
    ehyb_layout_free(L);
    return rc;
}

int ehyb_layout_build(const matrixCOO *m, const ehyb_layout_opts *opts_in, ehyb_layout **out)
{
    if (!m || !out) return ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_build: NULL argument");
    ehyb_layout_opts o;
    if (opts_in) o = *opts_in;
    else { memset(&o, 0, sizeof o); o.er_fill = 0.5; }
    if (o.W <= 0) o.W = m->vectorCacheSize;
    if (o.ctasPerPart <= 0) o.ctasPerPart = m->kernelPerPart > 0 ? m->kernelPerPart : 1;
    const int n = m->dimension;
    if (n <= 0 || !m->rowIdx || !m->partBoundary) return ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_build: matrix not reordered");
    int64_t *ptr = (int64_t *)malloc(((size_t)n + 1) * sizeof(int64_t));
    if (!ptr) return ehyb_fail(EHYB_ERR_NOMEM, "layout: out of memory");
    for (int i = 0; i <= n; ++i) ptr[i] = m->rowIdx[i];
    int rc = EHYB_OK;
    for (int64_t e = 0; e + 1 < m->totalNum && rc == EHYB_OK; ++e)
        if (m->I[e] > m->I[e + 1]) rc = ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_build: entries are not row-sorted");
    if (rc == EHYB_OK) rc = ehyb_layout_build_csr(n, ptr, m->J, m->V, m->nParts, m->partBoundary, &o, out);
    free(ptr);
    return rc;
}

/* ---------------------------------------------------------------------------------- */
/* streamed build: the matrix arrives as chunks of consecutive partitions               */
/* ---------------------------------------------------------------------------------- */

/*
 * The one-shot build holds the whole permuted matrix (12 B per entry) next to the layout
 * (~10.5 B per entry) - and its callers the generator output and the unpermuted copy before that:
 * ~70 B per entry at the peak, which is what kept BASELINE.json config 5 (27-point 512^3, 3.6 G
 * entries) out of reach.  The builder takes the permuted rows a few partitions at a time
 * (CSR, absolute local column numbers: the permutation is known up front) and appends their
 * slices to one blob, so the peak is the layout itself plus one chunk.  The result is the layout
 * the one-shot build produces for the same rows, options and er_fill >= 0, byte for byte
 * (tests/test_stream_build.py); er_fill < 0 chooses per chunk instead of per matrix.
 */
struct ehyb_layout_builder {
    ehyb_layout_opts opts;
    int64_t n, ncols;
    ehyb_layout *L;
    int64_t rowsDone;
    int64_t partsCap, slicesCap, blobCap, cacheCap, ovfCap;
    int keepRowInfo;
};

void ehyb_layout_builder_abort(ehyb_layout_builder *B)
{
    if (!B) return;
    ehyb_layout_free(B->L);
    free(B);
}

int ehyb_layout_builder_begin(int64_t n, const ehyb_layout_opts *opts, ehyb_layout_builder **out)
{
    if (!opts || !out || n <= 0 || n > INT_MAX || opts->W <= 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_builder_begin: bad argument");
    ehyb_layout_builder *B = (ehyb_layout_builder *)calloc(1, sizeof *B);
    if (!B) return ehyb_fail(EHYB_ERR_NOMEM, "layout builder: out of memory");
    B->opts = *opts;
    B->n = n;
    B->ncols = opts->ncols > 0 ? opts->ncols : n;
    B->opts.ncols = B->ncols;
    B->opts.min_coverage = -1.0; /* the all-overflow fallback is a whole-matrix decision */
    B->L = (ehyb_layout *)calloc(1, sizeof(ehyb_layout));
    if (!B->L) { free(B); return ehyb_fail(EHYB_ERR_NOMEM, "layout builder: out of memory"); }
    /* per-row bookkeeping (20 B per row) only serves the de-interleave and the cache file */
    B->keepRowInfo = n <= ((int64_t)64 << 20);
    if (B->keepRowInfo) {
        B->L->rowEll = (int32_t *)calloc((size_t)n, sizeof(int32_t));
        B->L->rowRemIn = (int32_t *)calloc((size_t)n, sizeof(int32_t));
        B->L->rowCached = (int32_t *)calloc((size_t)n, sizeof(int32_t));
        B->L->ovfPtr = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
        if (!B->L->rowEll || !B->L->rowRemIn || !B->L->rowCached || !B->L->ovfPtr) { ehyb_layout_builder_abort(B); return ehyb_fail(EHYB_ERR_NOMEM, "layout builder: out of memory"); }
    }
    *out = B;
    return EHYB_OK;
}

static int grow(void **p, int64_t *cap, int64_t need, size_t elem)
{
    if (need <= *cap) return 0;
    int64_t nc = *cap + *cap / 2 + 1024;
    if (nc < need) nc = need;
    void *q = realloc(*p, (size_t)nc * elem);
    if (!q) return -1;
    *p = q;
    *cap = nc;
    return 0;
}

int ehyb_layout_builder_add(ehyb_layout_builder *B, int nParts, const int32_t *partBoundary, const int64_t *rowPtr, const int32_t *col,
                            const double *val)
{
    if (!B || !partBoundary || !rowPtr || !col || !val || nParts <= 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_builder_add: bad argument");
    if (partBoundary[0] != B->rowsDone) return ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_builder_add: the chunk starts at row %d, %lld rows were added so far", partBoundary[0], (long long)B->rowsDone);
    const int64_t nRows = (int64_t)partBoundary[nParts] - partBoundary[0];
    if (nRows <= 0 || B->rowsDone + nRows > B->n) return ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_builder_add: bad row range");
    ehyb_layout *L = B->L, *C = NULL;
    ehyb_layout_view *v = &L->v;
    /* the blob: allocated after the first chunk from its bytes per row (+3 %), zero-filled lazily
     * by calloc; later chunks are written in place while they fit */
    blob_target bt = {NULL, 0, 0};
    if (L->blob && v->blobBytes < B->blobCap) { bt.dst = L->blob + v->blobBytes; bt.cap = B->blobCap - v->blobBytes; }
    int rc;
    if (nRows == B->n) rc = layout_build_impl(B->n, 0, nRows, rowPtr, col, val, nParts, partBoundary, &B->opts, 0, NULL, &C);
    else rc = layout_build_impl(B->n, B->rowsDone, nRows, rowPtr, col, val, nParts, partBoundary, &B->opts, 0, &bt, &C);
    if (rc) return rc;
    const ehyb_layout_view *cv = &C->v;
    rc = EHYB_ERR_NOMEM;
    if (grow((void **)&L->parts, &B->partsCap, (int64_t)v->nParts + nParts, sizeof(ehyb_part_desc))) goto done;
    if (grow((void **)&L->slices, &B->slicesCap, (int64_t)v->nSlices + cv->nSlices, sizeof(ehyb_slice_desc))) goto done;
    if (grow((void **)&L->cacheCols, &B->cacheCap, v->cacheTotal + cv->cacheTotal + 1, sizeof(int32_t))) goto done;
    if (cv->nOverflow > 0) {
        int64_t c1 = B->ovfCap, c2 = B->ovfCap, c3 = B->ovfCap;
        if (grow((void **)&L->ovfRow, &c1, v->nOverflow + cv->nOverflow, sizeof(int32_t)) || grow((void **)&L->ovfCol, &c2, v->nOverflow + cv->nOverflow, sizeof(int32_t)) ||
            grow((void **)&L->ovfVal, &c3, v->nOverflow + cv->nOverflow, sizeof(double))) goto done;
        B->ovfCap = c1 < c2 ? (c1 < c3 ? c1 : c3) : (c2 < c3 ? c2 : c3);
    }
    if ((v->blobBytes + cv->blobBytes) / 256 > (int64_t)UINT32_MAX || (int64_t)v->nSlices + cv->nSlices > INT_MAX || v->cacheTotal + cv->cacheTotal > INT_MAX ||
        v->nOverflow + cv->nOverflow > INT_MAX) { rc = ehyb_fail(EHYB_ERR_LIMIT, "layout builder: a 32-bit offset would overflow"); goto done; }
    if (!bt.used) {
        if (v->blobBytes + cv->blobBytes > B->blobCap) {
            /* first chunk: size the blob for the whole matrix; later: the estimate was short */
            const double perRow = (double)(v->blobBytes + cv->blobBytes) / (double)(B->rowsDone + nRows);
            int64_t want = (int64_t)(perRow * (double)B->n * 1.03) + ((int64_t)1 << 20);
            if (want < v->blobBytes + cv->blobBytes) want = v->blobBytes + cv->blobBytes;
            want = (want + 255) & ~(int64_t)255;
            unsigned char *nb = (unsigned char *)calloc((size_t)want, 1);
            if (!nb) { rc = ehyb_fail(EHYB_ERR_NOMEM, "layout builder: out of memory (%lld bytes)", (long long)want); goto done; }
            if (v->blobBytes) memcpy(nb, L->blob, (size_t)v->blobBytes);
            free(L->blob);
            L->blob = nb;
            B->blobCap = want;
        }
        memcpy(L->blob + v->blobBytes, cv->blob, (size_t)cv->blobBytes);
    }
    for (int p = 0; p < nParts; ++p) {
        ehyb_part_desc d = cv->parts[p];
        d.sliceStart += v->nSlices; d.sliceEnd += v->nSlices;
        d.cacheStart += (int32_t)v->cacheTotal;
        L->parts[v->nParts + p] = d;
    }
    const uint32_t off0 = (uint32_t)(v->blobBytes / 256);
    for (int s2 = 0; s2 < cv->nSlices; ++s2) {
        ehyb_slice_desc d = cv->slices[s2];
        d.off256 += off0;
        L->slices[v->nSlices + s2] = d;
    }
    if (cv->cacheTotal) memcpy(L->cacheCols + v->cacheTotal, cv->cacheCols, (size_t)cv->cacheTotal * sizeof(int32_t));
    if (cv->nOverflow) {
        memcpy(L->ovfRow + v->nOverflow, cv->ovfRow, (size_t)cv->nOverflow * sizeof(int32_t));
        memcpy(L->ovfCol + v->nOverflow, cv->ovfCol, (size_t)cv->nOverflow * sizeof(int32_t));
        memcpy(L->ovfVal + v->nOverflow, cv->ovfVal, (size_t)cv->nOverflow * sizeof(double));
    }
    if (B->keepRowInfo) {
        memcpy(L->rowEll + B->rowsDone, C->rowEll, (size_t)nRows * sizeof(int32_t));
        memcpy(L->rowRemIn + B->rowsDone, C->rowRemIn, (size_t)nRows * sizeof(int32_t));
        memcpy(L->rowCached + B->rowsDone, C->rowCached, (size_t)nRows * sizeof(int32_t));
        for (int64_t r = 0; r <= nRows; ++r) L->ovfPtr[B->rowsDone + r] = v->nOverflow + C->ovfPtr[r];
    }
    v->nParts += nParts; v->nSlices += cv->nSlices; v->blobBytes += cv->blobBytes; v->cacheTotal += cv->cacheTotal; v->nOverflow += cv->nOverflow;
    if (cv->cacheMax > v->cacheMax) v->cacheMax = cv->cacheMax;
    v->nnz += cv->nnz; v->nnzEll += cv->nnzEll; v->nnzRemInSlice += cv->nnzRemInSlice; v->nnzOverflow += cv->nnzOverflow;
    v->padEll += cv->padEll; v->padRem += cv->padRem; v->nLongRows += cv->nLongRows;
    v->haloInOverflow = cv->haloInOverflow;
    v->ctasPerPart = cv->ctasPerPart;
    B->rowsDone += nRows;
    rc = EHYB_OK;
done:
    if (rc == EHYB_ERR_NOMEM) ehyb_fail(EHYB_ERR_NOMEM, "layout builder: out of memory");
    ehyb_layout_free(C);
    return rc;
}

int ehyb_layout_builder_finish(ehyb_layout_builder *B, ehyb_layout **out)
{
    if (!B || !out) return ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_builder_finish: NULL argument");
    if (B->rowsDone != B->n) return ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_builder_finish: %lld of %lld rows were added", (long long)B->rowsDone, (long long)B->n);
    ehyb_layout *L = B->L;
    ehyb_layout_view *v = &L->v;
    if (B->blobCap > v->blobBytes + ((int64_t)1 << 20)) { /* give the unused tail of the estimate back */
        unsigned char *nb = (unsigned char *)realloc(L->blob, (size_t)(v->blobBytes ? v->blobBytes : 256));
        if (nb) L->blob = nb;
    }
    if (!L->ovfRow) { /* the device session expects valid (if empty) arrays */
        L->ovfRow = (int32_t *)malloc(sizeof(int32_t)); L->ovfCol = (int32_t *)malloc(sizeof(int32_t)); L->ovfVal = (double *)malloc(sizeof(double));
    }
    if (!L->cacheCols) L->cacheCols = (int32_t *)malloc(sizeof(int32_t));
    v->n = B->n; v->ncols = B->ncols; v->W = B->opts.W;
    v->parts = L->parts; v->slices = L->slices; v->blob = L->blob;
    v->ovfRow = L->ovfRow; v->ovfCol = L->ovfCol; v->ovfVal = L->ovfVal; v->cacheCols = L->cacheCols;
    v->algBytes = 8 * v->nnz + 2 * v->nnzEll + 4 * (v->nnz - v->nnzEll) + 8 * v->ncols + 8 * v->n;
    v->formatBytes = v->blobBytes + (int64_t)v->nSlices * (int64_t)sizeof(ehyb_slice_desc) + (int64_t)v->nParts * (int64_t)sizeof(ehyb_part_desc) +
                     v->nOverflow * 16 + v->cacheTotal * 4;
    *out = L;
    free(B);
    return EHYB_OK;
}

/* ---------------------------------------------------------------------------------- */
/* de-interleave: tuned layout -> reference layout                                     */
/* ---------------------------------------------------------------------------------- */

int ehyb_layout_to_reference(const ehyb_layout *L, matrixEHYB *out, int *sizeBlockELL, int *sizeER)
{
    if (!L || !out || !sizeBlockELL || !sizeER) return ehyb_fail(EHYB_ERR_ARG, "ehyb_layout_to_reference: NULL argument");
    const ehyb_layout_view *v = &L->v;
    if (v->nnz > INT_MAX || v->ncols != v->n) return ehyb_fail(EHYB_ERR_LIMIT, "reference layout needs nnz < 2^31 and no halo columns");
    if (!L->rowEll || !L->ovfPtr) return ehyb_fail(EHYB_ERR_LIMIT, "this layout was streamed without its per-row bookkeeping (more than 64 M rows)");
    const int n = (int)v->n, P = v->nParts;
    const int64_t nnz = v->nnz;
    matrixCOO c;
    memset(&c, 0, sizeof c);
    c.dimension = n; c.totalNum = (int)nnz; c.nParts = P; c.vectorCacheSize = (uint16_t)v->W; c.kernelPerPart = (int16_t)v->ctasPerPart;
    c.rowIdx = (int *)malloc(((size_t)n + 1) * sizeof(int));
    c.numInRow = (int *)malloc((size_t)n * sizeof(int));
    c.numInRow2 = (int *)calloc((size_t)n, sizeof(int));
    c.I = (int *)malloc((size_t)(nnz ? nnz : 1) * sizeof(int));
    c.J = (int *)malloc((size_t)(nnz ? nnz : 1) * sizeof(int));
    c.V = (double *)malloc((size_t)(nnz ? nnz : 1) * sizeof(double));
    c.partBoundary = (int *)malloc(((size_t)P + 1) * sizeof(int));
    int rc = EHYB_OK;
    if (!c.rowIdx || !c.numInRow || !c.numInRow2 || !c.I || !c.J || !c.V || !c.partBoundary) {
        rc = ehyb_fail(EHYB_ERR_NOMEM, "de-interleave: out of memory");
        goto done;
    }
    for (int p = 0; p < P; ++p) c.partBoundary[p] = v->parts[p].rowStart;
    c.partBoundary[P] = n;
    c.rowIdx[0] = 0;
    for (int r = 0; r < n; ++r) {
        c.numInRow[r] = L->rowEll[r] + L->rowRemIn[r] + (int)(L->ovfPtr[r + 1] - L->ovfPtr[r]);
        c.rowIdx[r + 1] = c.rowIdx[r] + c.numInRow[r];
    }
    if (c.rowIdx[n] != nnz) { rc = ehyb_fail(EHYB_ERR_ARG, "de-interleave: count mismatch"); goto done; }
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < P; ++p) {
        const int ps = v->parts[p].rowStart, pe = v->parts[p].rowEnd;
        for (int s = v->parts[p].sliceStart; s < v->parts[p].sliceEnd; ++s) {
            const int r0 = ps + (s - v->parts[p].sliceStart) * SR;
            const int r1 = r0 + SR < pe ? r0 + SR : pe;
            const int w = v->slices[s].w, wr = v->slices[s].wr;
            const unsigned char *base = v->blob + (int64_t)v->slices[s].off256 * 256;
            const double *ev = (const double *)base;
            const uint16_t *ec = (const uint16_t *)(base + reg_ell_col(w));
            const double *rv = (const double *)(base + reg_rem_val(w));
            const uint16_t *rcol = (const uint16_t *)(base + reg_rem_col(w, wr));
            const int32_t *cl = v->cacheCols + v->parts[p].cacheStart;
            for (int r = r0; r < r1; ++r) {
                const int t = r - r0, lane = t % 32, h = t / 32;
                int dst = c.rowIdx[r];
                /* ELL entries first, then the remainder: the reference layout depends on the
                 * order inside each class only */
                for (int k = 0; k < L->rowEll[r]; ++k, ++dst) {
                    c.I[dst] = r;
                    c.J[dst] = ps + ec[(((int64_t)(k / 4) * 32 + lane) * 2 + h) * 4 + k % 4];
                    c.V[dst] = ev[((int64_t)k * 32 + lane) * 2 + h];
                }
                for (int k = 0; k < L->rowRemIn[r]; ++k, ++dst) {
                    c.I[dst] = r;
                    c.J[dst] = cl[rcol[(((int64_t)(k / 4) * 32 + lane) * 2 + h) * 4 + k % 4]];
                    c.V[dst] = rv[((int64_t)k * 32 + lane) * 2 + h];
                }
                for (int64_t o = L->ovfPtr[r]; o < L->ovfPtr[r + 1]; ++o, ++dst) {
                    c.I[dst] = v->ovfRow[o];
                    c.J[dst] = v->ovfCol[o];
                    c.V[dst] = v->ovfVal[o];
                }
                int inw = 0;
                for (int e = c.rowIdx[r]; e < c.rowIdx[r + 1]; ++e) inw += (c.J[e] >= ps && c.J[e] < ps + v->W);
                c.numInRow2[r] = inw;
            }
        }
    }
    rc = ehyb_convert_reference_layout(&c, out, sizeBlockELL, sizeER, 1);
    if (rc == EHYB_OK) {
        out->partBoundary = c.partBoundary; /* owned by the caller now */
        c.partBoundary = NULL;
    }
done:
    free(c.rowIdx); free(c.numInRow); free(c.numInRow2); free(c.I); free(c.J); free(c.V); free(c.partBoundary);
    return rc;
}

/* ---------------------------------------------------------------------------------- */
/* serialisation hooks (cache.c)                                                       */
/* ---------------------------------------------------------------------------------- */

/* every array of a layout, in the order ehyb_layout_import expects them */
int ehyb_layout_export_arrays(const ehyb_layout *L, const void **arrays, int64_t *bytes, int max)
{
    const ehyb_layout_view *v = &L->v;
    const void *a[] = {L->parts, L->slices, L->blob, L->ovfRow, L->ovfCol, L->ovfVal, L->cacheCols,
                       L->rowEll, L->rowRemIn, L->rowCached, L->ovfPtr};
    const int64_t b[] = {(int64_t)sizeof(ehyb_part_desc) * v->nParts, (int64_t)sizeof(ehyb_slice_desc) * v->nSlices, v->blobBytes,
                         4 * v->nOverflow, 4 * v->nOverflow, 8 * v->nOverflow, 4 * v->cacheTotal,
                         4 * v->n, 4 * v->n, 4 * v->n, 8 * (v->n + 1)};
    const int na = (int)(sizeof a / sizeof a[0]);
    if (max < na) return 0;
    for (int i = 0; i < na; ++i) { arrays[i] = a[i]; bytes[i] = b[i]; }
    return na;
}

/* builds a layout around malloc'd arrays (ownership passes to the layout) after checking that
 * they are consistent with the scalars: a corrupt cache must not reach the device */
int ehyb_layout_import(const ehyb_layout_view *sc, void *const *arrays, ehyb_layout **out)
{
    ehyb_layout *L = (ehyb_layout *)calloc(1, sizeof *L);
    if (!L) return ehyb_fail(EHYB_ERR_NOMEM, "layout import: out of memory");
    const ehyb_part_desc *parts = (const ehyb_part_desc *)arrays[0];
    const ehyb_slice_desc *slices = (const ehyb_slice_desc *)arrays[1];
    const int32_t *cacheCols = (const int32_t *)arrays[6];
    int bad = sc->W <= 0 || sc->W > 65536 || sc->ncols < sc->n;
    for (int p = 0; p < sc->nParts && !bad; ++p) {
        const ehyb_part_desc *d = &parts[p];
        bad = d->rowStart < 0 || d->rowEnd < d->rowStart || d->rowEnd > sc->n || d->sliceStart < 0 || d->sliceEnd < d->sliceStart ||
              d->sliceEnd > sc->nSlices || d->cacheStart < 0 || d->cacheCount < 0 || (int64_t)d->cacheStart + d->cacheCount > sc->cacheTotal ||
              d->cacheCount > sc->cacheMax || (p > 0 && d->rowStart != parts[p - 1].rowEnd);
    }
    for (int s = 0; s < sc->nSlices && !bad; ++s)
        bad = (int64_t)slices[s].off256 * 256 + slice_bytes(slices[s].w, slices[s].wr) > sc->blobBytes;
    /* every row of a partition lives in one of its slices, and every 16-bit index stays inside
     * what the kernels stage in shared memory: the x window (rows [rowStart, min(rowStart+W, n)))
     * for ELL columns, the partition's cache list for remainder columns.  A bad index would be an
     * out-of-bounds shared-memory read on the device, a short slice range rows of y never written. */
    const unsigned char *blob = (const unsigned char *)arrays[2];
    int badIdx = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(| : badIdx)
    for (int p = 0; p < sc->nParts; ++p) {
        if (bad) continue;
        const ehyb_part_desc *d = &parts[p];
        if (d->sliceEnd - d->sliceStart != (d->rowEnd - d->rowStart + SR - 1) / SR) { badIdx = 1; continue; }
        const int64_t winLen = (int64_t)d->rowStart + sc->W < sc->n ? sc->W : sc->n - d->rowStart;
        for (int s = d->sliceStart; s < d->sliceEnd; ++s) {
            const int w = slices[s].w, wr = slices[s].wr;
            const unsigned char *base = blob + (int64_t)slices[s].off256 * 256;
            const uint16_t *ec = (const uint16_t *)(base + reg_ell_col(w));
            const uint16_t *rcol = (const uint16_t *)(base + reg_rem_col(w, wr));
            if (w > 0 && winLen <= 0) { badIdx = 1; break; }
            if (wr > 0 && d->cacheCount <= 0) { badIdx = 1; break; }
            const int64_t nE = (int64_t)((w + 3) / 4) * 256, nR = (int64_t)((wr + 3) / 4) * 256;
            for (int64_t i = 0; i < nE; ++i) badIdx |= ec[i] >= winLen;
            for (int64_t i = 0; i < nR; ++i) badIdx |= rcol[i] >= d->cacheCount;
        }
    }
    bad |= badIdx;
    for (int64_t i = 0; i < sc->cacheTotal && !bad; ++i) bad = cacheCols[i] < 0 || cacheCols[i] >= sc->ncols;
    const int32_t *ovfRow = (const int32_t *)arrays[3], *ovfCol = (const int32_t *)arrays[4];
    for (int64_t i = 0; i < sc->nOverflow && !bad; ++i)
        bad = ovfRow[i] < 0 || ovfRow[i] >= sc->n || ovfCol[i] < 0 || ovfCol[i] >= sc->ncols || (i > 0 && ovfRow[i] < ovfRow[i - 1]);
    if (bad) { free(L); return ehyb_fail(EHYB_ERR_IO, "layout import: arrays are inconsistent with the header"); }
    L->parts = (ehyb_part_desc *)arrays[0]; L->slices = (ehyb_slice_desc *)arrays[1]; L->blob = (unsigned char *)arrays[2];
    L->ovfRow = (int32_t *)arrays[3]; L->ovfCol = (int32_t *)arrays[4]; L->ovfVal = (double *)arrays[5];
    L->cacheCols = (int32_t *)arrays[6]; L->rowEll = (int32_t *)arrays[7]; L->rowRemIn = (int32_t *)arrays[8];
    L->rowCached = (int32_t *)arrays[9]; L->ovfPtr = (int64_t *)arrays[10];
    L->v = *sc;
    L->v.parts = L->parts; L->v.slices = L->slices; L->v.blob = L->blob;
    L->v.ovfRow = L->ovfRow; L->v.ovfCol = L->ovfCol; L->v.ovfVal = L->ovfVal; L->v.cacheCols = L->cacheCols;
    *out = L;
    return EHYB_OK;
}
