/*
 * hierpart.c -- deterministic AND parallel partition stage.
 *
 * The reference partitions with one mt-metis call: one thread on the symmetric path
 * (reordering.c:274; 8.6 s of the 10.5 s host time at config 2), six threads on the general path
 * (reordering.c:120), where the result changes from run to run.  mt-metis is a pinned binary
 * (no source), so the way to get both properties is around it:
 *
 *   level 0  the rows are contracted in blocks of g consecutive rows (<= 32 768 blocks; vertex
 *            weight = rows, edge weight = matrix entries between two blocks) and the coarse graph
 *            is cut into T pieces by ONE single-threaded mt-metis call (milliseconds);
 *   level 1  every piece gets its share of the nparts partitions (largest remainders of
 *            rows_t / n) and is partitioned on its own - the subgraph induced by its rows - by a
 *            single-threaded mt-metis process; the T processes run at the same time.
 *
 * Every call is single-threaded with the library's fixed seed, so the partition vector is the
 * same on every run, and the wall time is that of the largest piece.  The price is the cut between
 * the pieces, chosen on the coarse graph (config 2, T = 8: +3 % remainder entries).  The result is
 * a partition vector like any other: ehyb_reorder_with_partition() takes it from there, and the
 * parity of everything downstream "given the same partition" is untouched.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "common.h"

static int g_pieces = 0;

void ehyb_set_partition_pieces(int pieces) { g_pieces = pieces > 0 ? pieces : 0; }

int ehyb_get_partition_pieces(void)
{
    const char *s = getenv("EHYB_PARTITION_PIECES");
    return s && s[0] ? atoi(s) : g_pieces;
}

int ehyb_partition_graph_hier(uint32_t n, const uint32_t *xadj, const uint32_t *adjncy, uint32_t nparts, int pieces, uint32_t *where)
{
    if (!xadj || !adjncy || !where || nparts == 0 || n == 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_partition_graph_hier: bad argument");
    int T = pieces;
    if (T > (int)nparts) T = (int)nparts;
    if (T <= 1 || n < 65536u) return ehyb_partition_graph(n, xadj, adjncy, nparts, 1, where);
    const uint32_t g = (n + 32767u) / 32768u;     /* rows per block */
    const uint32_t nb = (n + g - 1) / g;
    int rc = EHYB_OK;
    uint32_t *cx = (uint32_t *)calloc((size_t)nb + 1, sizeof(uint32_t)), *cadj = NULL, *piece = (uint32_t *)malloc((size_t)nb * sizeof(uint32_t));
    int32_t *cvw = (int32_t *)malloc((size_t)nb * sizeof(int32_t)), *caw = NULL;
    int32_t *local = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    int64_t *rowsOf = (int64_t *)calloc((size_t)T, sizeof(int64_t));
    uint32_t *partsOf = (uint32_t *)calloc((size_t)T + 1, sizeof(uint32_t));
    if (!cx || !piece || !cvw || !local || !rowsOf || !partsOf) { rc = ehyb_fail(EHYB_ERR_NOMEM, "partition: out of memory"); goto done; }

    /* ---- level 0: the block graph (two passes: degrees, then edges) ---- */
    for (int pass = 0; pass < 2 && rc == EHYB_OK; ++pass) {
        int oom = 0;
#pragma omp parallel reduction(| : oom)
        {
            int32_t *acc = (int32_t *)calloc((size_t)nb, sizeof(int32_t));
            uint32_t *touched = (uint32_t *)malloc((size_t)nb * sizeof(uint32_t));
            if (!acc || !touched) oom = 1;
#pragma omp for schedule(dynamic, 64)
            for (uint32_t b = 0; b < nb; ++b) {
                if (oom) continue;
                const uint32_t r0 = b * g, r1 = r0 + g < n ? r0 + g : n;
                uint32_t nt = 0;
                for (uint32_t e = xadj[r0]; e < xadj[r1]; ++e) {
                    const uint32_t c = adjncy[e] / g;
                    if (c == b) continue;
                    if (acc[c]++ == 0) touched[nt++] = c;
                }
                if (pass == 0) {
                    cx[b + 1] = nt;
                    cvw[b] = (int32_t)(r1 - r0);
                } else {
                    /* ascending neighbour order: the graph handed to mt-metis must not depend on thread timing */
                    for (uint32_t i = 1; i < nt; ++i) { /* insertion sort: a block has few neighbours */
                        const uint32_t v = touched[i];
                        uint32_t j = i;
                        while (j > 0 && touched[j - 1] > v) { touched[j] = touched[j - 1]; --j; }
                        touched[j] = v;
                    }
                    for (uint32_t i = 0; i < nt; ++i) { cadj[cx[b] + i] = touched[i]; caw[cx[b] + i] = acc[touched[i]]; }
                }
                for (uint32_t i = 0; i < nt; ++i) acc[touched[i]] = 0;
            }
            free(acc); free(touched);
        }
        if (oom) { rc = ehyb_fail(EHYB_ERR_NOMEM, "partition: out of memory"); break; }
        if (pass == 0) {
            for (uint32_t b = 0; b < nb; ++b) cx[b + 1] += cx[b];
            cadj = (uint32_t *)malloc((size_t)(cx[nb] ? cx[nb] : 1) * sizeof(uint32_t));
            caw = (int32_t *)malloc((size_t)(cx[nb] ? cx[nb] : 1) * sizeof(int32_t));
            if (!cadj || !caw) rc = ehyb_fail(EHYB_ERR_NOMEM, "partition: out of memory");
        }
    }
    if (rc) goto done;
    /* an unsymmetric block graph (general matrices hand in A + A^T, so this does not happen) would be
     * rejected by the partitioner; the weights a->b and b->a agree because the pattern is symmetric */
    rc = ehyb_partition_graph_weighted(nb, cx, cadj, cvw, caw, (uint32_t)T, 1, 1.001f, piece);
    if (rc) goto done;

    /* ---- level 1: the pieces, each with its share of the partitions ---- */
    for (uint32_t b = 0; b < nb; ++b) rowsOf[piece[b]] += cvw[b];
    {
        /* largest remainders, at least one partition per non-empty piece */
        double *frac = (double *)malloc((size_t)T * sizeof(double));
        if (!frac) { rc = ehyb_fail(EHYB_ERR_NOMEM, "partition: out of memory"); goto done; }
        uint32_t given = 0;
        for (int t = 0; t < T; ++t) {
            const double share = (double)nparts * (double)rowsOf[t] / (double)n;
            uint32_t k = (uint32_t)share;
            if (rowsOf[t] > 0 && k == 0) k = 1;
            partsOf[t + 1] = k;
            frac[t] = share - (double)(uint32_t)share;
            given += k;
        }
        while (given < nparts) {
            int best = -1;
            for (int t = 0; t < T; ++t)
                if (rowsOf[t] > 0 && (best < 0 || frac[t] > frac[best])) best = t;
            if (best < 0) break;
            partsOf[best + 1] += 1; frac[best] = -1.0; given += 1;
        }
        while (given > nparts) { /* (only after the "at least one" rule) take from the piece with the most */
            int big = 0;
            for (int t = 1; t < T; ++t)
                if (partsOf[t + 1] > partsOf[big + 1]) big = t;
            partsOf[big + 1] -= 1; given -= 1;
        }
        free(frac);
        for (int t = 0; t < T; ++t) partsOf[t + 1] += partsOf[t];
    }
    {
        int64_t *cnt = (int64_t *)calloc((size_t)T, sizeof(int64_t));
        if (!cnt) { rc = ehyb_fail(EHYB_ERR_NOMEM, "partition: out of memory"); goto done; }
        for (uint32_t i = 0; i < n; ++i) local[i] = (int32_t)cnt[piece[i / g]]++;
        free(cnt);
    }
    int fail = 0;
    char msg[512] = "";
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        const uint32_t nt = (uint32_t)rowsOf[t], kt = partsOf[t + 1] - partsOf[t];
        if (nt == 0) continue;
        uint32_t *sx = (uint32_t *)malloc(((size_t)nt + 1) * sizeof(uint32_t));
        uint32_t *rows = (uint32_t *)malloc((size_t)nt * sizeof(uint32_t));
        uint32_t *w = (uint32_t *)malloc((size_t)nt * sizeof(uint32_t));
        uint32_t *sa = NULL;
        int bad = !sx || !rows || !w;
        if (!bad) {
            uint64_t m = 0;
            uint32_t k = 0;
            for (uint32_t i = 0; i < n; ++i)
                if (piece[i / g] == (uint32_t)t) {
                    rows[k++] = i;
                    for (uint32_t e = xadj[i]; e < xadj[i + 1]; ++e) m += piece[adjncy[e] / g] == (uint32_t)t;
                }
            sa = (uint32_t *)malloc((size_t)(m ? m : 1) * sizeof(uint32_t));
            bad = !sa;
        }
        if (!bad) {
            uint32_t o = 0;
            for (uint32_t k = 0; k < nt; ++k) {
                sx[k] = o;
                const uint32_t i = rows[k];
                for (uint32_t e = xadj[i]; e < xadj[i + 1]; ++e)
                    if (piece[adjncy[e] / g] == (uint32_t)t) sa[o++] = (uint32_t)local[adjncy[e]];
            }
            sx[nt] = o;
            int r2 = ehyb_partition_graph_process(nt, sx, sa, kt, w);
            if (r2) {
#pragma omp critical
                { fail = r2; snprintf(msg, sizeof msg, "%s", ehyb_last_error()); }
            } else {
                for (uint32_t k = 0; k < nt; ++k) where[rows[k]] = partsOf[t] + w[k];
            }
        } else {
#pragma omp critical
            { fail = EHYB_ERR_NOMEM; snprintf(msg, sizeof msg, "partition: out of memory"); }
        }
        free(sx); free(rows); free(w); free(sa);
    }
    if (fail) rc = ehyb_fail(fail, "%s", msg);
done:
    free(cx); free(cadj); free(piece); free(cvw); free(caw); free(local); free(rowsOf); free(partsOf);
    return rc;
}
