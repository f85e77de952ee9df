/*
 * reorder.c -- host partition + reorder stage.
 *
 * Same results as the reference's reordering.c (bit-exact permutation, partition map and
 * permuted arrays for the same partition vector), organised differently:
 *   - the graph build, the partitioner call and the permutation are separate entry points,
 *     so a partition vector can be injected (tests, structured grids, multi-GPU blocks);
 *   - the in-partition row order is an explicit key (same-partition entry count descending,
 *     original index ascending) instead of relying on what qsort does with a comparator
 *     that returns 0 on ties (reference Partition.h:17-24);
 *   - partitions are sorted, and row-sorted inputs are scattered, in parallel (OpenMP).
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "common.h"
#include "reordering.h"

/* ---------------------------------------------------------------------------------- */
/* graph                                                                               */
/* ---------------------------------------------------------------------------------- */

int ehyb_build_graph(const matrixCOO *m, int symmetric, uint32_t **xadj_out, uint32_t **adj_out)
{
    if (!m || !xadj_out || !adj_out || m->dimension <= 0 || m->totalNum < 0)
        return ehyb_fail(EHYB_ERR_ARG, "ehyb_build_graph: bad argument");
    const int n = m->dimension;
    const int64_t nnz = m->totalNum;
    for (int64_t e = 0; e < nnz; ++e)
        if ((unsigned)m->I[e] >= (unsigned)n || (unsigned)m->J[e] >= (unsigned)n)
            return ehyb_fail(EHYB_ERR_ARG, "entry %lld (%d,%d) outside a %d x %d matrix", (long long)e,
                             m->I[e], m->J[e], n, n);
    uint32_t *xadj = (uint32_t *)malloc(((size_t)n + 1) * sizeof(uint32_t));
    uint32_t *adj = NULL;
    if (!xadj) return ehyb_fail(EHYB_ERR_NOMEM, "graph: out of memory");
    if (symmetric) {
        /* The pattern itself, self loops included: reordering.c:239-264. */
        adj = (uint32_t *)malloc((size_t)(nnz ? nnz : 1) * sizeof(uint32_t));
        if (!adj) { free(xadj); return ehyb_fail(EHYB_ERR_NOMEM, "graph: out of memory"); }
        for (int i = 0; i <= n; ++i) xadj[i] = (uint32_t)m->rowIdx[i];
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < nnz; ++e) adj[e] = (uint32_t)m->J[e];
    } else {
        /* Pattern of A + A^T; (i,j) lands in row i and, off the diagonal, in row j, in
         * entry order; duplicates are kept: reordering.c:56-89. */
        if (2 * nnz > (int64_t)UINT32_MAX) { free(xadj); return ehyb_fail(EHYB_ERR_LIMIT, "graph larger than 2^32 edges"); }
        uint32_t *fill = (uint32_t *)calloc((size_t)n + 1, sizeof(uint32_t));
        adj = (uint32_t *)malloc((size_t)(nnz > 0 ? 2 * nnz : 1) * sizeof(uint32_t));
        if (!fill || !adj) { free(xadj); free(fill); free(adj); return ehyb_fail(EHYB_ERR_NOMEM, "graph: out of memory"); }
        for (int64_t e = 0; e < nnz; ++e) {
            fill[m->I[e]] += 1;
            if (m->I[e] != m->J[e]) fill[m->J[e]] += 1;
        }
        xadj[0] = 0;
        for (int i = 0; i < n; ++i) { xadj[i + 1] = xadj[i] + fill[i]; fill[i] = 0; }
        for (int64_t e = 0; e < nnz; ++e) {
            const int i = m->I[e], j = m->J[e];
            adj[xadj[i] + fill[i]++] = (uint32_t)j;
            if (i != j) adj[xadj[j] + fill[j]++] = (uint32_t)i;
        }
        free(fill);
    }
    *xadj_out = xadj;
    *adj_out = adj;
    return EHYB_OK;
}

/* ---------------------------------------------------------------------------------- */
/* permutation from a partition vector                                                 */
/* ---------------------------------------------------------------------------------- */

typedef struct { uint32_t key; uint32_t idx; } row_key;

static int row_key_cmp(const void *a, const void *b)
{
    const row_key *x = (const row_key *)a, *y = (const row_key *)b;
    if (x->key != y->key) return x->key > y->key ? -1 : 1; /* more same-partition entries first */
    return x->idx < y->idx ? -1 : (x->idx > y->idx);       /* then original order */
}

int ehyb_reorder_with_partition(matrixCOO *m, const uint32_t *partVec)
{
    return ehyb_reorder_core(m, partVec, m ? m->dimension : 0);
}

/* Rows and the first n columns are permuted; columns in [n, ncols) are the halo of a local
 * block of a distributed matrix: they keep their index, belong to no partition and to no
 * window.  ncols == n is the reference's case. */
int ehyb_reorder_core(matrixCOO *m, const uint32_t *partVec, int ncols)
{
    if (!m || !partVec || m->dimension <= 0 || m->nParts <= 0 || ncols < m->dimension)
        return ehyb_fail(EHYB_ERR_ARG, "ehyb_reorder_with_partition: bad argument");
    const int n = m->dimension, P = m->nParts;
    const int64_t nnz = m->totalNum;
    const int W = m->vectorCacheSize;
    for (int i = 0; i < n; ++i)
        if (partVec[i] >= (uint32_t)P) return ehyb_fail(EHYB_ERR_ARG, "partVec[%d] = %u >= nParts %d", i, partVec[i], P);

    int rc = EHYB_OK;
    int *bound = (int *)calloc((size_t)(P + 1 > n ? P + 1 : n) + 1, sizeof(int));
    int *cursor = (int *)calloc((size_t)P, sizeof(int));
    row_key *rows = (row_key *)malloc((size_t)n * sizeof(row_key));
    uint32_t *same = (uint32_t *)calloc((size_t)n, sizeof(uint32_t));
    int *newLen = (int *)calloc((size_t)n, sizeof(int));
    int *newI = (int *)malloc((size_t)(nnz ? nnz : 1) * sizeof(int));
    int *newJ = (int *)malloc((size_t)(nnz ? nnz : 1) * sizeof(int));
    double *newV = (double *)malloc((size_t)(nnz ? nnz : 1) * sizeof(double));
    int64_t *newPtr = (int64_t *)malloc(((size_t)n + 1) * sizeof(int64_t));
    if (!bound || !cursor || !rows || !same || !newLen || !newI || !newJ || !newV || !newPtr) {
        rc = ehyb_fail(EHYB_ERR_NOMEM, "reorder: out of memory");
        goto done;
    }
    int *perm = m->reorderList;

    /* partition sizes -> boundaries (reordering.c:301-307, :319-321) */
    for (int i = 0; i < n; ++i) bound[partVec[i] + 1] += 1;
    for (int p = 0; p < P; ++p) bound[p + 1] += bound[p];

    /* same-partition entry count per row (reordering.c:327-331); entries may be unsorted */
    int sorted = 1;
    for (int64_t e = 0; e < nnz; ++e) {
        const int i = m->I[e], j = m->J[e];
        if ((unsigned)i >= (unsigned)n || (unsigned)j >= (unsigned)ncols) {
            rc = ehyb_fail(EHYB_ERR_ARG, "entry %lld outside the matrix", (long long)e);
            goto done;
        }
        same[i] += j < n && partVec[i] == partVec[j];
        if (e && m->I[e - 1] > i) sorted = 0;
    }
    if (sorted)
        for (int i = 0; i < n && sorted; ++i)
            if (m->rowIdx[i] > m->rowIdx[i + 1] || m->rowIdx[n] != nnz) sorted = 0;

    /* rows bucketed by partition in original order, then sorted inside each partition */
    for (int i = 0; i < n; ++i) {
        row_key *r = &rows[bound[partVec[i]] + cursor[partVec[i]]++];
        r->key = same[i];
        r->idx = (uint32_t)i;
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < P; ++p)
        qsort(rows + bound[p], (size_t)(bound[p + 1] - bound[p]), sizeof(row_key), row_key_cmp);
    for (int r = 0; r < n; ++r) perm[rows[r].idx] = r;

    /* new row lengths and pointers (reordering.c:335-345) */
    for (int i = 0; i < n; ++i) newLen[perm[i]] += m->rowIdx[i + 1] - m->rowIdx[i];
    newPtr[0] = 0;
    for (int r = 0; r < n; ++r) newPtr[r + 1] = newPtr[r] + newLen[r];
    if (newPtr[n] != nnz) {
        rc = ehyb_fail(EHYB_ERR_ARG, "rowIdx sums to %lld entries, totalNum is %lld", (long long)newPtr[n], (long long)nnz);
        goto done;
    }

    /* scatter (reordering.c:348-362): per-row entry order is preserved */
    memset(m->numInRow, 0, (size_t)n * sizeof(int));
    memset(m->numInRow2, 0, (size_t)n * sizeof(int));
    if (sorted) {
        int mismatch = 0;
#pragma omp parallel for schedule(static) reduction(| : mismatch)
        for (int i = 0; i < n; ++i) {
            const int r = perm[i];
            const int ps = bound[partVec[i]], pe = ps + W;
            int64_t dst = newPtr[r];
            int inWin = 0;
            for (int e = m->rowIdx[i]; e < m->rowIdx[i + 1]; ++e, ++dst) {
                const int c = m->J[e] < n ? perm[m->J[e]] : m->J[e];
                mismatch |= m->I[e] != i;
                newI[dst] = r;
                newJ[dst] = c;
                newV[dst] = m->V[e];
                inWin += (c >= ps && c < pe);
            }
            m->numInRow[r] = m->rowIdx[i + 1] - m->rowIdx[i];
            m->numInRow2[r] = inWin;
        }
        if (mismatch) {
            rc = ehyb_fail(EHYB_ERR_ARG, "rowIdx does not describe the row-sorted entries");
            goto done;
        }
    } else {
        for (int64_t e = 0; e < nnz; ++e) {
            const int i = m->I[e];
            const int r = perm[i], c = m->J[e] < n ? perm[m->J[e]] : m->J[e];
            const int64_t dst = newPtr[r] + m->numInRow[r]++;
            const int ps = bound[partVec[i]];
            newI[dst] = r;
            newJ[dst] = c;
            newV[dst] = m->V[e];
            m->numInRow2[r] += (c >= ps && c < ps + W);
        }
    }
    for (int r = 0; r <= n; ++r) m->rowIdx[r] = (int)newPtr[r];

    /* ownership as in the reference (reordering.c:363-369): the caller's I/J/V are released
     * and replaced; partBoundary is re-allocated (the reference leaks the old one, B-16) */
    free(m->I); free(m->J); free(m->V);
    m->I = newI; m->J = newJ; m->V = newV;
    newI = NULL; newJ = NULL; newV = NULL;
    free(m->partBoundary);
    m->partBoundary = bound;
    bound = NULL;

done:
    free(bound); free(cursor); free(rows); free(same); free(newLen);
    free(newI); free(newJ); free(newV); free(newPtr);
    return rc;
}

int ehyb_reorder(matrixCOO *m, int symmetric)
{
    if (!m) return ehyb_fail(EHYB_ERR_ARG, "ehyb_reorder: NULL matrix");
    uint32_t *xadj = NULL, *adj = NULL;
    printf("nParts is %d\n", m->nParts); /* reordering.c:237 */
    int rc = ehyb_build_graph(m, symmetric, &xadj, &adj);
    if (rc) return rc;
    uint32_t *where = (uint32_t *)calloc((size_t)m->dimension, sizeof(uint32_t));
    if (!where) { free(xadj); free(adj); return ehyb_fail(EHYB_ERR_NOMEM, "reorder: out of memory"); }
    struct timeval t0, t1;
    gettimeofday(&t0, NULL);
    printf("start k-way partition\n"); /* reordering.c:279 */
    /* 1 thread on the symmetric path (reordering.c:274), 6 on the other (:120) */
    const int pieces = ehyb_get_partition_pieces();
    if (pieces > 1) /* deterministic and parallel: hierpart.c (also on the general path: no threaded mt-metis call) */
        rc = ehyb_partition_graph_hier((uint32_t)m->dimension, xadj, adj, (uint32_t)m->nParts, pieces, where);
    else
        rc = ehyb_partition_graph((uint32_t)m->dimension, xadj, adj, (uint32_t)m->nParts, symmetric ? 1u : 6u, where);
    free(xadj);
    free(adj);
    if (rc == EHYB_OK) {
        printf("partition finished\n");
        gettimeofday(&t1, NULL);
        printf("partition time is %ld us\n", (t1.tv_sec * 1000000 + t1.tv_usec) - (t0.tv_sec * 1000000 + t0.tv_usec));
        rc = ehyb_reorder_with_partition(m, where);
    }
    free(where);
    return rc;
}

void matrixReorder(matrixCOO *m)
{
    if (ehyb_reorder(m, 1)) ehyb_die("matrixReorder");
}

void matrixReorder_unsym(matrixCOO *m)
{
    if (ehyb_reorder(m, 0)) ehyb_die("matrixReorder_unsym");
}

void vectorReorder(const int n, const double *v_in, double *v_rodr, const int *list)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) v_rodr[list[i]] = v_in[i];
}

void vectorRecover(const int n, const double *v_rodr, double *v, const int *list)
{
    /* the reference builds the inverse list first (reordering.c:387-389); a gather through
     * the forward list gives the same vector */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) v[i] = v_rodr[list[i]];
}
