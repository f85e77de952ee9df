/*
 * matgen.c -- synthetic matrices of the BASELINE.json configs and the
 * reader's expansion into the matrixCOO the pipeline starts from.
 *
 * The symmetric generators emit exactly the entry sequence a Matrix Market file of the matrix
 * holds (lower triangle, column-major), so that generating in memory and reading the .mtx
 * give the same row-sorted COO - including the per-row entry order, which fixes the fp64
 * summation order and the EHYB index arrays.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.h"

typedef struct {
    int *i, *j;
    double *v;
    int64_t count, cap;
} entries;

static int push(entries *E, int r, int c, double v)
{
    if (E->count == E->cap) {
        int64_t cap = E->cap ? E->cap * 2 : (1 << 20);
        int *ni = (int *)realloc(E->i, (size_t)cap * sizeof(int));
        if (ni) E->i = ni;
        int *nj = (int *)realloc(E->j, (size_t)cap * sizeof(int));
        if (nj) E->j = nj;
        double *nv = (double *)realloc(E->v, (size_t)cap * sizeof(double));
        if (nv) E->v = nv;
        if (!ni || !nj || !nv) return -1;
        E->cap = cap;
    }
    E->i[E->count] = r; E->j[E->count] = c; E->v[E->count] = v;
    E->count++;
    return 0;
}

int ehyb_gen_lower(ehyb_gen_kind kind, int nx, int ny, int nz, int *n_out, int64_t *count, int **li, int **lj, double **lv)
{
    if (!n_out || !count || !li || !lj || !lv || nx <= 0 || ny <= 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_gen_lower: bad argument");
    if (kind == EHYB_GEN_LAPLACE2D) nz = 1;
    if (nz <= 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_gen_lower: bad nz");
    const int64_t nodes = (int64_t)nx * ny * nz;
    const int dof = kind == EHYB_GEN_ELASTICITY ? 3 : 1;
    if (nodes * dof > INT_MAX) return ehyb_fail(EHYB_ERR_LIMIT, "generator: more than 2^31 rows");
    entries E = {0};
    int bad = 0;
    if (kind == EHYB_GEN_LAPLACE2D) {
        /* per column c: (c,c)=4, (c+1,c)=-1 unless c ends a grid line, (c+nx,c)=-1 unless last line */
        for (int64_t c = 0; c < nodes && !bad; ++c) {
            bad |= push(&E, (int)c, (int)c, 4.0);
            if (c % nx != nx - 1) bad |= push(&E, (int)c + 1, (int)c, -1.0);
            if (c + nx < nodes) bad |= push(&E, (int)c + nx, (int)c, -1.0);
        }
    } else if (kind == EHYB_GEN_STENCIL27 || kind == EHYB_GEN_ELASTICITY) {
        /* per column node: dz in {0,1}, dy, dx in {-1,0,1}, keep row >= column */
        for (int64_t cn = 0; cn < nodes && !bad; ++cn) {
            const int x = (int)(cn % nx), y = (int)((cn / nx) % ny), z = (int)(cn / ((int64_t)nx * ny));
            int64_t nb[18];
            int m = 0;
            for (int dz = 0; dz <= 1; ++dz)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (x + dx < 0 || x + dx >= nx || y + dy < 0 || y + dy >= ny || z + dz >= nz) continue;
                        const int64_t rn = cn + (int64_t)dz * nx * ny + (int64_t)dy * nx + dx;
                        if (rn >= cn) nb[m++] = rn;
                    }
            if (dof == 1) {
                for (int k = 0; k < m && !bad; ++k) bad |= push(&E, (int)nb[k], (int)cn, nb[k] == cn ? 26.0 : -1.0);
            } else {
                /* 3x3 blocks: column 3*cn+dc; rows 3*rn+dr, dr = 0..2; value 100 on the diagonal,
                 * else -(1 + 0.25*((7r+13c) mod 4)) */
                for (int dc = 0; dc < 3 && !bad; ++dc) {
                    const int64_t c = 3 * cn + dc;
                    for (int k = 0; k < m && !bad; ++k)
                        for (int dr = 0; dr < 3; ++dr) {
                            const int64_t r = 3 * nb[k] + dr;
                            if (r < c) continue;
                            bad |= push(&E, (int)r, (int)c, r == c ? 100.0 : -(1.0 + 0.25 * (double)((7 * r + 13 * c) % 4)));
                        }
                }
            }
        }
    } else {
        return ehyb_fail(EHYB_ERR_ARG, "ehyb_gen_lower: unknown kind %d", (int)kind);
    }
    if (bad) { free(E.i); free(E.j); free(E.v); return ehyb_fail(EHYB_ERR_NOMEM, "generator: out of memory"); }
    *n_out = (int)(nodes * dof);
    *count = E.count;
    *li = E.i; *lj = E.j; *lv = E.v;
    return EHYB_OK;
}

void ehyb_x_reference(int n, double *x)
{
    for (int i = 0; i < n; ++i) { /* solver_test.c:228-232 */
        srand((unsigned)i);
        x[i] = (double)(rand() % 200 - 100) / 1000;
    }
}

static int coo_alloc(matrixCOO *m, int n, int64_t nnz)
{
    memset(m, 0, sizeof *m);
    if (nnz > INT_MAX) return ehyb_fail(EHYB_ERR_LIMIT, "matrixCOO holds at most 2^31-1 entries (got %lld)", (long long)nnz);
    m->dimension = n;
    m->totalNum = (int)nnz;
    m->kernelPerPart = 1;
    m->rowIdx = (int *)calloc((size_t)n + 1, sizeof(int));
    m->numInRow = (int *)calloc((size_t)n, sizeof(int));
    m->numInRow2 = (int *)calloc((size_t)n, sizeof(int));
    m->I = (int *)malloc((size_t)(nnz ? nnz : 1) * sizeof(int));
    m->J = (int *)malloc((size_t)(nnz ? nnz : 1) * sizeof(int));
    m->V = (double *)malloc((size_t)(nnz ? nnz : 1) * sizeof(double));
    m->diag = (double *)calloc((size_t)n, sizeof(double));
    m->partBoundary = (int *)calloc((size_t)n + 1, sizeof(int));
    m->reorderList = (int *)calloc((size_t)n, sizeof(int));
    if (!m->rowIdx || !m->numInRow || !m->numInRow2 || !m->I || !m->J || !m->V || !m->diag || !m->partBoundary || !m->reorderList) {
        ehyb_coo_free(m);
        return ehyb_fail(EHYB_ERR_NOMEM, "matrixCOO: out of memory");
    }
    return EHYB_OK;
}

void ehyb_coo_free(matrixCOO *m)
{
    if (!m) return;
    free(m->rowIdx); free(m->numInRow); free(m->numInRow2); free(m->I); free(m->J); free(m->V);
    free(m->diag); free(m->partBoundary); free(m->reorderList);
    memset(m, 0, sizeof *m);
}

int ehyb_coo_from_lower(int n, int64_t count, const int *li, const int *lj, const double *lv, matrixCOO *out,
                        const double *x, double *yg)
{
    if (!out || !li || !lj || !lv || n <= 0 || count < 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_coo_from_lower: bad argument");
    if (yg && !x) return ehyb_fail(EHYB_ERR_ARG, "ehyb_coo_from_lower: golden y needs x");
    int64_t ndiag = 0;
    for (int64_t k = 0; k < count; ++k) {
        if ((unsigned)li[k] >= (unsigned)n || (unsigned)lj[k] >= (unsigned)n)
            return ehyb_fail(EHYB_ERR_ARG, "entry %lld outside the matrix", (long long)k);
        ndiag += li[k] == lj[k];
    }
    /* the reference assumes every diagonal entry is present (totalNum = 2*lower - n, B-12);
     * counting the diagonal entries gives the same number then and stays right otherwise */
    const int64_t nnz = 2 * count - ndiag;
    int rc = coo_alloc(out, n, nnz);
    if (rc) return rc;
    int *fill = out->numInRow;
    for (int64_t k = 0; k < count; ++k) { /* solver_test.c:196-206 */
        fill[li[k]] += 1;
        if (li[k] != lj[k]) fill[lj[k]] += 1;
    }
    int maxCol = 0;
    for (int i = 0; i < n; ++i) { /* :214-222 */
        if (fill[i] > maxCol) maxCol = fill[i];
        out->rowIdx[i + 1] = out->rowIdx[i] + fill[i];
        fill[i] = 0;
    }
    out->maxCol = maxCol;
    if (yg) memset(yg, 0, (size_t)n * sizeof(double));
    for (int64_t k = 0; k < count; ++k) { /* :235-260: file order, both triangles */
        const int r = li[k], c = lj[k];
        const double v = lv[k];
        int dst = out->rowIdx[r] + fill[r]++;
        out->I[dst] = r; out->J[dst] = c; out->V[dst] = v;
        if (yg) yg[r] += v * x[c];
        if (r != c) {
            dst = out->rowIdx[c] + fill[c]++;
            out->I[dst] = c; out->J[dst] = r; out->V[dst] = v;
            if (yg) yg[c] += v * x[r];
        } else {
            out->diag[r] = v;
        }
    }
    return EHYB_OK;
}

int ehyb_coo_from_general(int n, int64_t count, const int *fi, const int *fj, const double *fv, matrixCOO *out,
                          const double *x, double *yg)
{
    if (!out || !fi || !fj || !fv || n <= 0 || count < 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_coo_from_general: bad argument");
    if (yg && !x) return ehyb_fail(EHYB_ERR_ARG, "ehyb_coo_from_general: golden y needs x");
    for (int64_t k = 0; k < count; ++k)
        if ((unsigned)fi[k] >= (unsigned)n || (unsigned)fj[k] >= (unsigned)n)
            return ehyb_fail(EHYB_ERR_ARG, "entry %lld outside the matrix", (long long)k);
    int rc = coo_alloc(out, n, count);
    if (rc) return rc;
    if (yg) memset(yg, 0, (size_t)n * sizeof(double));
    for (int64_t k = 0; k < count; ++k) { /* solver_test.c:96-103: entries stay in file order */
        out->I[k] = fi[k]; out->J[k] = fj[k]; out->V[k] = fv[k];
        out->numInRow[fi[k]] += 1;
        if (yg) yg[fi[k]] += fv[k] * x[fj[k]];
        if (fi[k] == fj[k]) out->diag[fi[k]] = fv[k];
    }
    int maxCol = 0;
    for (int i = 0; i < n; ++i) { /* :111-121 */
        if (out->numInRow[i] > maxCol) maxCol = out->numInRow[i];
        out->rowIdx[i + 1] = out->rowIdx[i] + out->numInRow[i];
    }
    out->maxCol = maxCol;
    return EHYB_OK;
}

/* ---------------------------------------------------------------------------------- */
/* R-MAT                                                                               */
/* ---------------------------------------------------------------------------------- */

static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

typedef struct { uint64_t key; double v; } kv;

/* LSD radix sort on the low `bits` bits of key, 11 bits per pass; stable */
static int radix_sort_kv(kv *a, int64_t m, int bits)
{
    kv *tmp = (kv *)malloc((size_t)(m ? m : 1) * sizeof(kv));
    if (!tmp) return -1;
    kv *src = a, *dst = tmp;
    for (int shift = 0; shift < bits; shift += 11) {
        int64_t cnt[2049];
        memset(cnt, 0, sizeof cnt);
        for (int64_t i = 0; i < m; ++i) cnt[((src[i].key >> shift) & 2047) + 1]++;
        for (int k = 0; k < 2048; ++k) cnt[k + 1] += cnt[k];
        for (int64_t i = 0; i < m; ++i) dst[cnt[(src[i].key >> shift) & 2047]++] = src[i];
        kv *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, (size_t)m * sizeof(kv));
    free(tmp);
    return 0;
}

int ehyb_gen_rmat(int scale, int edge_factor, uint64_t seed, int add_diagonal, int *n_out, int64_t *count,
                  int **fi, int **fj, double **fv)
{
    if (scale < 1 || scale > 30 || edge_factor < 1 || !n_out || !count || !fi || !fj || !fv)
        return ehyb_fail(EHYB_ERR_ARG, "ehyb_gen_rmat: bad argument");
    const int n = 1 << scale;
    const int64_t m0 = (int64_t)n * edge_factor, m = m0 + (add_diagonal ? n : 0);
    const double a = 0.57, b = 0.19, c = 0.19;
    kv *E = (kv *)malloc((size_t)m * sizeof(kv));
    if (!E) return ehyb_fail(EHYB_ERR_NOMEM, "rmat: out of memory");
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < m0; ++e) {
        uint64_t r = 0, cc = 0;
        for (int l = 0; l < scale; ++l) {
            const uint64_t h = mix64(mix64((uint64_t)e * 64 + (uint64_t)l) ^ (seed * 0x51ED27ULL));
            const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
            const int qb = u >= a && u < a + b, qc = u >= a + b && u < a + b + c, qd = u >= a + b + c;
            r = (r << 1) | (uint64_t)(qc | qd);
            cc = (cc << 1) | (uint64_t)(qb | qd);
        }
        const uint64_t hv = mix64((uint64_t)e ^ ((seed + 77) * 0x2545F491ULL));
        E[e].key = r * (uint64_t)n + cc;
        E[e].v = (double)(hv >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
    for (int64_t d = 0; d < (add_diagonal ? n : 0); ++d) {
        E[m0 + d].key = (uint64_t)d * (uint64_t)n + (uint64_t)d;
        E[m0 + d].v = 4.0;
    }
    if (radix_sort_kv(E, m, 2 * scale)) { free(E); return ehyb_fail(EHYB_ERR_NOMEM, "rmat: out of memory"); }
    int64_t u = 0;
    for (int64_t e = 0; e < m; ++e) { /* duplicates summed in generation order */
        if (u && E[u - 1].key == E[e].key) E[u - 1].v += E[e].v;
        else E[u++] = E[e];
    }
    int *I = (int *)malloc((size_t)u * sizeof(int)), *J = (int *)malloc((size_t)u * sizeof(int));
    double *V = (double *)malloc((size_t)u * sizeof(double));
    if (!I || !J || !V) { free(E); free(I); free(J); free(V); return ehyb_fail(EHYB_ERR_NOMEM, "rmat: out of memory"); }
    for (int64_t e = 0; e < u; ++e) {
        I[e] = (int)(E[e].key / (uint64_t)n);
        J[e] = (int)(E[e].key % (uint64_t)n);
        V[e] = E[e].v;
    }
    free(E);
    *n_out = n; *count = u; *fi = I; *fj = J; *fv = V;
    return EHYB_OK;
}
