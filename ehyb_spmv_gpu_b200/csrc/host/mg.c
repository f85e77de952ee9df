/*
 * mg.c -- host side of the multi-GPU layer (no reference counterpart).
 *
 * The global matrix is distributed by contiguous blocks of rows (the level-1 partition: metis
 * row blocks or grid slabs, made contiguous by the caller's level-1 permutation); rank g owns
 * rows [rowStarts[g], rowStarts[g+1]) and the matching entries of x and y.  Every rank runs one
 * process with one GPU.  This file turns a rank's block of rows (global column indices) into
 *   - the halo: the sorted list of global columns the block references outside its own
 *     range, grouped by owner -> what the rank receives each product, and where
 *     (local column index n_local + position in the list);
 *   - the local operator: columns renumbered [own | halo], rows/own columns permuted by the
 *     level-2 (per-GPU) partition exactly like the single-GPU path (ehyb_reorder_core), and
 *     built into the tuned layout - for the NCCL exchange with every halo entry in the overflow
 *     list, so that the main kernel never depends on the exchange; for the peer-memory
 *     exchange with halo columns in the partitions' remainder caches;
 *   - the send list: permuted local indices of the x entries each peer asked for.
 * The per-product data path (peer-memory push / NCCL send/recv, launches) is the "multi-GPU"
 * part of cuda/ehyb_device.cu.
 * Who needs what is exchanged between the ranks by the caller (any transport: the Python
 * front end uses torch.distributed, gloo on CPU and nccl on GPU), which keeps this file free
 * of communication and testable without GPUs.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.h"

#include "mg_internal.h"

static int cmp_i64(const void *a, const void *b)
{
    const int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return x < y ? -1 : x > y;
}

void ehyb_mg_local_free(ehyb_mg_local *L)
{
    if (!L) return;
    free(L->rowStarts); free(L->rowPtr); free(L->col); free(L->val);
    free(L->haloGlobal); free(L->recvCount); free(L->sendCount); free(L->sendGlobal); free(L->sendIdx);
    if (L->finished && !L->streamed) ehyb_coo_free(&L->coo);
    free(L->perm);
    ehyb_layout_free(L->layout);
    free(L);
}

int ehyb_mg_local_build(int rank, int nranks, const int64_t *rowStarts, const int64_t *rowPtr, const int64_t *colGlobal,
                        const double *val, ehyb_mg_local **out)
{
    if (!rowStarts || !rowPtr || !colGlobal || !val || !out || nranks <= 0 || rank < 0 || rank >= nranks)
        return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_local_build: bad argument");
    const int64_t r0 = rowStarts[rank], r1 = rowStarts[rank + 1], n = r1 - r0, N = rowStarts[nranks];
    if (n <= 0 || n > INT_MAX) return ehyb_fail(EHYB_ERR_ARG, "rank %d owns %lld rows", rank, (long long)n);
    const int64_t nnz = rowPtr[n];
    if (nnz > INT_MAX) return ehyb_fail(EHYB_ERR_LIMIT, "a local block holds at most 2^31-1 entries (got %lld)", (long long)nnz);
    ehyb_mg_local *L = (ehyb_mg_local *)calloc(1, sizeof *L);
    if (!L) return ehyb_fail(EHYB_ERR_NOMEM, "mg: out of memory");
    int rc = EHYB_OK;
    int64_t *ext = NULL;
    L->rank = rank; L->nranks = nranks; L->n = n; L->nnz = nnz;
    L->rowStarts = (int64_t *)malloc(((size_t)nranks + 1) * sizeof(int64_t));
    L->rowPtr = (int64_t *)malloc(((size_t)n + 1) * sizeof(int64_t));
    L->col = (int32_t *)malloc((size_t)(nnz ? nnz : 1) * sizeof(int32_t));
    L->val = (double *)malloc((size_t)(nnz ? nnz : 1) * sizeof(double));
    L->recvCount = (int64_t *)calloc((size_t)nranks, sizeof(int64_t));
    L->sendCount = (int64_t *)calloc((size_t)nranks, sizeof(int64_t));
    if (!L->rowStarts || !L->rowPtr || !L->col || !L->val || !L->recvCount || !L->sendCount) { rc = ehyb_fail(EHYB_ERR_NOMEM, "mg: out of memory"); goto fail; }
    memcpy(L->rowStarts, rowStarts, ((size_t)nranks + 1) * sizeof(int64_t));
    memcpy(L->rowPtr, rowPtr, ((size_t)n + 1) * sizeof(int64_t));
    memcpy(L->val, val, (size_t)nnz * sizeof(double));

    /* halo = sorted unique external columns */
    int64_t nExt = 0;
    for (int64_t e = 0; e < nnz; ++e) {
        const int64_t c = colGlobal[e];
        if (c < 0 || c >= N) { rc = ehyb_fail(EHYB_ERR_ARG, "column %lld outside the global matrix", (long long)c); goto fail; }
        nExt += (c < r0 || c >= r1);
    }
    ext = (int64_t *)malloc((size_t)(nExt ? nExt : 1) * sizeof(int64_t));
    if (!ext) { rc = ehyb_fail(EHYB_ERR_NOMEM, "mg: out of memory"); goto fail; }
    nExt = 0;
    for (int64_t e = 0; e < nnz; ++e)
        if (colGlobal[e] < r0 || colGlobal[e] >= r1) ext[nExt++] = colGlobal[e];
    qsort(ext, (size_t)nExt, sizeof(int64_t), cmp_i64);
    int64_t nHalo = 0;
    for (int64_t i = 0; i < nExt; ++i)
        if (i == 0 || ext[i] != ext[i - 1]) ext[nHalo++] = ext[i];
    if (n + nHalo > INT_MAX) { rc = ehyb_fail(EHYB_ERR_LIMIT, "local block + halo exceed 2^31 columns"); goto fail; }
    L->nHalo = nHalo;
    L->haloGlobal = ext;
    ext = NULL;
    int owner = 0;
    for (int64_t i = 0; i < nHalo; ++i) { /* the list is sorted, so owners come in order */
        while (L->haloGlobal[i] >= rowStarts[owner + 1]) ++owner;
        L->recvCount[owner] += 1;
    }
    /* renumber columns: own -> global - r0, halo -> n + position */
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < nnz; ++e) {
        const int64_t c = colGlobal[e];
        if (c >= r0 && c < r1) {
            L->col[e] = (int32_t)(c - r0);
        } else {
            int64_t lo = 0, hi = nHalo - 1;
            while (lo < hi) {
                const int64_t mid = (lo + hi) / 2;
                if (L->haloGlobal[mid] < c) lo = mid + 1; else hi = mid;
            }
            L->col[e] = (int32_t)(n + lo);
        }
    }
    *out = L;
    return EHYB_OK;
fail:
    free(ext);
    ehyb_mg_local_free(L);
    return rc;
}

int ehyb_mg_local_halo(const ehyb_mg_local *L, int64_t *nHalo, const int64_t **haloGlobal, const int64_t **recvCount)
{
    if (!L) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_local_halo: NULL");
    if (nHalo) *nHalo = L->nHalo;
    if (haloGlobal) *haloGlobal = L->haloGlobal;
    if (recvCount) *recvCount = L->recvCount;
    return EHYB_OK;
}

int ehyb_mg_local_set_send(ehyb_mg_local *L, const int64_t *sendCount, const int64_t *sendGlobal)
{
    if (!L || !sendCount) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_local_set_send: NULL");
    int64_t tot = 0;
    for (int g = 0; g < L->nranks; ++g) {
        if (sendCount[g] < 0 || (g == L->rank && sendCount[g] != 0)) return ehyb_fail(EHYB_ERR_ARG, "bad send count for peer %d", g);
        tot += sendCount[g];
    }
    if (tot > INT_MAX) return ehyb_fail(EHYB_ERR_LIMIT, "send list too long");
    const int64_t r0 = L->rowStarts[L->rank], r1 = L->rowStarts[L->rank + 1];
    for (int64_t i = 0; i < tot; ++i)
        if (!sendGlobal || sendGlobal[i] < r0 || sendGlobal[i] >= r1) return ehyb_fail(EHYB_ERR_ARG, "peer asked for row %lld, not owned by rank %d", sendGlobal ? (long long)sendGlobal[i] : -1LL, L->rank);
    free(L->sendGlobal);
    L->sendGlobal = (int64_t *)malloc((size_t)(tot ? tot : 1) * sizeof(int64_t));
    if (!L->sendGlobal) return ehyb_fail(EHYB_ERR_NOMEM, "mg: out of memory");
    memcpy(L->sendCount, sendCount, (size_t)L->nranks * sizeof(int64_t));
    if (tot) memcpy(L->sendGlobal, sendGlobal, (size_t)tot * sizeof(int64_t));
    L->nSend = tot;
    return ehyb_mg_local_update_send_idx(L);
}

int ehyb_mg_local_update_send_idx(ehyb_mg_local *L)
{
    const int32_t *perm = L->streamed ? L->perm : L->coo.reorderList;
    if (!L->finished || !perm) return EHYB_OK; /* set_send came first: finish() calls again */
    free(L->sendIdx);
    L->sendIdx = (int32_t *)malloc((size_t)(L->nSend ? L->nSend : 1) * sizeof(int32_t));
    if (!L->sendIdx) return ehyb_fail(EHYB_ERR_NOMEM, "mg: out of memory");
    const int64_t r0 = L->rowStarts[L->rank];
    for (int64_t i = 0; i < L->nSend; ++i) L->sendIdx[i] = perm[L->sendGlobal[i] - r0];
    return EHYB_OK;
}

/* Level-2: partition the own-column block (symmetrised pattern handled by the caller's choice
 * of partVec; NULL = contiguous blocks), permute, build the layout. */
int ehyb_mg_local_finish(ehyb_mg_local *L, int nParts, int W, int ctasPerPart, const uint32_t *partVec, double er_fill,
                         int exchange)
{
    if (!L || nParts <= 0 || W <= 0 || (exchange != EHYB_MG_NCCL && exchange != EHYB_MG_P2P))
        return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_local_finish: bad argument");
    if (L->finished) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_local_finish: already finished");
    const int n = (int)L->n;
    const int64_t nnz = L->nnz;
    matrixCOO *m = &L->coo;
    memset(m, 0, sizeof *m);
    m->dimension = n; m->totalNum = (int)nnz; m->nParts = nParts;
    m->vectorCacheSize = (uint16_t)(W > 65535 ? 65535 : W);
    m->kernelPerPart = (int16_t)(ctasPerPart > 0 ? ctasPerPart : 1);
    m->rowIdx = (int *)malloc(((size_t)n + 1) * sizeof(int));
    m->numInRow = (int *)calloc((size_t)n, sizeof(int));
    m->numInRow2 = (int *)calloc((size_t)n, sizeof(int));
    m->I = (int *)malloc((size_t)(nnz ? nnz : 1) * sizeof(int));
    m->J = (int *)malloc((size_t)(nnz ? nnz : 1) * sizeof(int));
    m->V = (double *)malloc((size_t)(nnz ? nnz : 1) * sizeof(double));
    m->diag = (double *)calloc((size_t)n, sizeof(double));
    m->partBoundary = (int *)calloc((size_t)n + 1, sizeof(int));
    m->reorderList = (int *)calloc((size_t)n, sizeof(int));
    uint32_t *pv = NULL;
    int rc = EHYB_OK;
    L->finished = 1; /* from here on the matrixCOO is released by ehyb_mg_local_free */
    if (!m->rowIdx || !m->numInRow || !m->numInRow2 || !m->I || !m->J || !m->V || !m->diag || !m->partBoundary || !m->reorderList)
        return ehyb_fail(EHYB_ERR_NOMEM, "mg: out of memory");
    int maxCol = 0;
    for (int r = 0; r < n; ++r) {
        m->rowIdx[r] = (int)L->rowPtr[r];
        m->numInRow[r] = (int)(L->rowPtr[r + 1] - L->rowPtr[r]);
        if (m->numInRow[r] > maxCol) maxCol = m->numInRow[r];
        for (int64_t e = L->rowPtr[r]; e < L->rowPtr[r + 1]; ++e) {
            m->I[e] = r;
            m->J[e] = L->col[e];
            m->V[e] = L->val[e];
        }
    }
    m->rowIdx[n] = (int)nnz;
    m->maxCol = maxCol;
    /* the natural-order copy is no longer needed */
    free(L->col); L->col = NULL;
    free(L->val); L->val = NULL;
    if (!partVec) {
        pv = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
        if (!pv) return ehyb_fail(EHYB_ERR_NOMEM, "mg: out of memory");
        ehyb_partition_blocks((uint32_t)n, (uint32_t)nParts, pv);
        partVec = pv;
    }
    rc = ehyb_reorder_core(m, partVec, (int)(n + L->nHalo));
    free(pv);
    if (rc) return rc;
    /* send list in permuted local numbering */
    rc = ehyb_mg_local_update_send_idx(L);
    if (rc) return rc;
    ehyb_layout_opts o;
    memset(&o, 0, sizeof o);
    o.W = W; o.ctasPerPart = ctasPerPart > 0 ? ctasPerPart : 1; o.er_fill = er_fill;
    o.ncols = n + L->nHalo;
    /* NCCL exchange: the halo arrives on another stream while the main kernel runs, so no halo
     * entry may sit in a slice; peer-memory exchange: halo columns are served from the
     * remainder cache like any other column outside the window */
    o.halo_in_overflow = exchange == EHYB_MG_NCCL;
    return ehyb_layout_build(m, &o, &L->layout);
}

int ehyb_mg_local_view(const ehyb_mg_local *L, const matrixCOO **coo, const ehyb_layout **layout, int64_t *nSend,
                       const int32_t **sendIdx, const int64_t **sendCount)
{
    if (!L || !L->finished) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_local_view: not finished");
    if (coo) *coo = L->streamed ? NULL : &L->coo;
    if (layout) *layout = L->layout;
    if (nSend) *nSend = L->nSend;
    if (sendIdx) *sendIdx = L->sendIdx;
    if (sendCount) *sendCount = L->sendCount;
    return EHYB_OK;
}

/* Own-column pattern of the block as a graph for the level-2 partitioner (symmetric input
 * assumed: the block's diagonal part is structurally symmetric when the matrix is). */
int ehyb_mg_local_graph(const ehyb_mg_local *L, uint32_t **xadj_out, uint32_t **adj_out)
{
    if (!L || !L->col || !xadj_out || !adj_out) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_local_graph: call before finish");
    const int n = (int)L->n;
    uint32_t *xadj = (uint32_t *)malloc(((size_t)n + 1) * sizeof(uint32_t));
    uint32_t *adj = (uint32_t *)malloc((size_t)(L->nnz ? L->nnz : 1) * sizeof(uint32_t));
    if (!xadj || !adj) { free(xadj); free(adj); return ehyb_fail(EHYB_ERR_NOMEM, "mg: out of memory"); }
    uint32_t k = 0;
    for (int r = 0; r < n; ++r) {
        xadj[r] = k;
        for (int64_t e = L->rowPtr[r]; e < L->rowPtr[r + 1]; ++e)
            if (L->col[e] < n) adj[k++] = (uint32_t)L->col[e];
    }
    xadj[n] = k;
    *xadj_out = xadj;
    *adj_out = adj;
    return EHYB_OK;
}

/* Rows [z0*nx*ny, z1*nx*ny) of the 27-point stencil on an nx x ny x nz grid (diag 26, off -1,
 * the BASELINE.json matrix), full rows with ascending global columns: the local block of a
 * z-slab decomposition, generated in place (config 5 never exists as a file). */
int ehyb_gen_stencil27_rows(int nx, int ny, int64_t nz, int64_t z0, int64_t z1, int64_t **rowPtr_out, int64_t **col_out,
                            double **val_out)
{
    if (nx <= 0 || ny <= 0 || nz <= 0 || z0 < 0 || z1 > nz || z0 >= z1 || !rowPtr_out || !col_out || !val_out)
        return ehyb_fail(EHYB_ERR_ARG, "ehyb_gen_stencil27_rows: bad argument");
    const int64_t plane = (int64_t)nx * ny, n = (z1 - z0) * plane;
    int64_t *rowPtr = (int64_t *)malloc(((size_t)n + 1) * sizeof(int64_t));
    if (!rowPtr) return ehyb_fail(EHYB_ERR_NOMEM, "generator: out of memory");
    rowPtr[0] = 0;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t g = z0 * plane + r;
        const int x = (int)(g % nx), y = (int)((g / nx) % ny);
        const int64_t z = g / plane;
        const int cx = 1 + (x > 0) + (x < nx - 1), cy = 1 + (y > 0) + (y < ny - 1), cz = 1 + (z > 0) + (z < nz - 1);
        rowPtr[r + 1] = rowPtr[r] + (int64_t)cx * cy * cz;
    }
    const int64_t nnz = rowPtr[n];
    int64_t *col = (int64_t *)malloc((size_t)nnz * sizeof(int64_t));
    double *val = (double *)malloc((size_t)nnz * sizeof(double));
    if (!col || !val) { free(rowPtr); free(col); free(val); return ehyb_fail(EHYB_ERR_NOMEM, "generator: out of memory"); }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        const int64_t g = z0 * plane + r;
        const int x = (int)(g % nx), y = (int)((g / nx) % ny);
        const int64_t z = g / plane;
        int64_t k = rowPtr[r];
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (x + dx < 0 || x + dx >= nx || y + dy < 0 || y + dy >= ny || z + dz < 0 || z + dz >= nz) continue;
                    const int64_t c = g + dz * plane + (int64_t)dy * nx + dx;
                    col[k] = c;
                    val[k] = c == g ? 26.0 : -1.0;
                    ++k;
                }
    }
    *rowPtr_out = rowPtr; *col_out = col; *val_out = val;
    return EHYB_OK;
}
