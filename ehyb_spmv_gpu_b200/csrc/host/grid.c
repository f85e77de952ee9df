/*
 * grid.c -- BASELINE.json config 5: the 27-point stencil on an nx x ny x nz grid (512^3: 134 M
 * rows, 3.6 G entries), distributed over the GPUs of one box.  No reference counterpart (the
 * reference is one GPU, 32-bit entry counts: SURVEY.md B-7, B-8); SURVEY.md section 8e.
 *
 * The matrix never exists as a COO.  The grid is cut into BRICKS of bx x by x bz cells:
 *
 *   level 1  bricks -> GPUs: any owner[] vector, in particular the k = G partition of the brick
 *            graph by the pinned mt-metis binary (ehyb_grid_brick_graph: vertex weight = cells,
 *            edge weight = matrix entries that couple the two bricks; the call of
 *            reordering.c:270-293 on the coarsened graph), or contiguous runs of bricks (slabs).
 *            Level-1 numbering: rank-major, then the rank's bricks in ascending brick order, then
 *            the cells of a brick in natural order - every rank owns a contiguous range of ids;
 *   level 2  a brick IS a partition of the EHYB format (its cells = the x window), rows sorted
 *            inside the brick by same-brick entry count descending, then natural order - the
 *            reference's rule (reordering.c:327-334) for the partition vector "brick of the row",
 *            evaluated in closed form instead of counting and qsort.
 *
 * ehyb_mg_grid_build() makes a rank's block in two passes over its bricks: (A) permutation and
 * halo (the level-1 ids of the cells other ranks own and this rank references), (B) the permuted
 * rows a few bricks at a time into the streamed layout builder.  Peak memory = the layout
 * (~10.8 B per entry) + 4 B per row + one chunk.  The result equals what the general path
 * (ehyb_mg_local_build + ehyb_mg_local_finish on ehyb_grid_rows, partition vector = brick) builds,
 * byte for byte (tests/test_stream_build.py).
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "mg_internal.h"

struct ehyb_grid_decomp {
    int nx, ny, nz, bx, by, bz, nbx, nby, nbz, nranks;
    int64_t nb;              /* bricks */
    int32_t *owner;          /* [nb] */
    int64_t *base;           /* [nb] level-1 id of the brick's first cell */
    int64_t *rowStarts;      /* [nranks + 1] */
    int32_t *rankBricks;     /* [nb] bricks grouped by owner, ascending */
    int64_t *rankBrickStart; /* [nranks + 1] */
};

void ehyb_grid_decomp_free(ehyb_grid_decomp *D)
{
    if (!D) return;
    free(D->owner); free(D->base); free(D->rowStarts); free(D->rankBricks); free(D->rankBrickStart);
    free(D);
}

static inline int imin(int a, int b) { return a < b ? a : b; }

typedef struct { int x0, y0, z0, sx, sy, sz; } brick_box;

static inline brick_box brick_extent(const ehyb_grid_decomp *D, int64_t b)
{
    brick_box e;
    const int ix = (int)(b % D->nbx), iy = (int)((b / D->nbx) % D->nby), iz = (int)(b / ((int64_t)D->nbx * D->nby));
    e.x0 = ix * D->bx; e.y0 = iy * D->by; e.z0 = iz * D->bz;
    e.sx = imin(D->bx, D->nx - e.x0); e.sy = imin(D->by, D->ny - e.y0); e.sz = imin(D->bz, D->nz - e.z0);
    return e;
}

static inline int64_t brick_of_cell(const ehyb_grid_decomp *D, int x, int y, int z)
{
    return ((int64_t)(z / D->bz) * D->nby + y / D->by) * D->nbx + x / D->bx;
}

/* level-1 id of grid cell (x, y, z) */
static inline int64_t level1_id(const ehyb_grid_decomp *D, int x, int y, int z)
{
    const int64_t b = brick_of_cell(D, x, y, z);
    const brick_box e = brick_extent(D, b);
    return D->base[b] + ((int64_t)(z - e.z0) * e.sy + (y - e.y0)) * e.sx + (x - e.x0);
}

int ehyb_grid_decomp_create(int nx, int ny, int nz, int bx, int by, int bz, int nranks, const uint32_t *owner, ehyb_grid_decomp **out)
{
    if (!out || nx <= 0 || ny <= 0 || nz <= 0 || bx <= 0 || by <= 0 || bz <= 0 || nranks <= 0)
        return ehyb_fail(EHYB_ERR_ARG, "ehyb_grid_decomp_create: bad argument");
    if ((int64_t)bx * by * bz > 65536) return ehyb_fail(EHYB_ERR_LIMIT, "a brick of %d x %d x %d cells exceeds the 16-bit window (65536)", bx, by, bz);
    ehyb_grid_decomp *D = (ehyb_grid_decomp *)calloc(1, sizeof *D);
    if (!D) return ehyb_fail(EHYB_ERR_NOMEM, "grid: out of memory");
    D->nx = nx; D->ny = ny; D->nz = nz; D->bx = bx; D->by = by; D->bz = bz; D->nranks = nranks;
    D->nbx = (nx + bx - 1) / bx; D->nby = (ny + by - 1) / by; D->nbz = (nz + bz - 1) / bz;
    D->nb = (int64_t)D->nbx * D->nby * D->nbz;
    if (D->nb > INT_MAX) { free(D); return ehyb_fail(EHYB_ERR_LIMIT, "too many bricks"); }
    D->owner = (int32_t *)malloc((size_t)D->nb * sizeof(int32_t));
    D->base = (int64_t *)malloc((size_t)D->nb * sizeof(int64_t));
    D->rowStarts = (int64_t *)calloc((size_t)nranks + 1, sizeof(int64_t));
    D->rankBricks = (int32_t *)malloc((size_t)D->nb * sizeof(int32_t));
    D->rankBrickStart = (int64_t *)calloc((size_t)nranks + 1, sizeof(int64_t));
    if (!D->owner || !D->base || !D->rowStarts || !D->rankBricks || !D->rankBrickStart) { ehyb_grid_decomp_free(D); return ehyb_fail(EHYB_ERR_NOMEM, "grid: out of memory"); }
    for (int64_t b = 0; b < D->nb; ++b) {
        /* default: contiguous runs of bricks in natural brick order (z-slabs of bricks) */
        const int64_t g = owner ? (int64_t)owner[b] : b * nranks / D->nb;
        if (g < 0 || g >= nranks) { ehyb_grid_decomp_free(D); return ehyb_fail(EHYB_ERR_ARG, "brick %lld is assigned to rank %lld of %d", (long long)b, (long long)g, nranks); }
        D->owner[b] = (int32_t)g;
        D->rankBrickStart[g + 1] += 1;
        const brick_box e = brick_extent(D, b);
        D->rowStarts[g + 1] += (int64_t)e.sx * e.sy * e.sz;
    }
    for (int g = 0; g < nranks; ++g) {
        D->rankBrickStart[g + 1] += D->rankBrickStart[g];
        D->rowStarts[g + 1] += D->rowStarts[g];
    }
    int64_t *cursor = (int64_t *)malloc((size_t)nranks * 2 * sizeof(int64_t));
    if (!cursor) { ehyb_grid_decomp_free(D); return ehyb_fail(EHYB_ERR_NOMEM, "grid: out of memory"); }
    for (int g = 0; g < nranks; ++g) { cursor[2 * g] = D->rankBrickStart[g]; cursor[2 * g + 1] = D->rowStarts[g]; }
    for (int64_t b = 0; b < D->nb; ++b) {
        const int g = D->owner[b];
        const brick_box e = brick_extent(D, b);
        D->rankBricks[cursor[2 * g]++] = (int32_t)b;
        D->base[b] = cursor[2 * g + 1];
        cursor[2 * g + 1] += (int64_t)e.sx * e.sy * e.sz;
    }
    free(cursor);
    *out = D;
    return EHYB_OK;
}

int ehyb_grid_decomp_info(const ehyb_grid_decomp *D, int64_t *nBricks, const int64_t **rowStarts, const int32_t **owner)
{
    if (!D) return ehyb_fail(EHYB_ERR_ARG, "ehyb_grid_decomp_info: NULL");
    if (nBricks) *nBricks = D->nb;
    if (rowStarts) *rowStarts = D->rowStarts;
    if (owner) *owner = D->owner;
    return EHYB_OK;
}

/* The brick graph for the level-1 partitioner: 26-neighbourhood, vwgt = cells of the brick,
 * adjwgt = matrix entries (one direction) between the two bricks: per axis 1 where the bricks
 * differ, 3 s - 2 where they share their extent s.  Arrays are malloc'd (ehyb_free_host). */
int ehyb_grid_brick_graph(int nx, int ny, int nz, int bx, int by, int bz, int64_t *nBricks, uint32_t **xadj_out, uint32_t **adj_out,
                          int32_t **vwgt_out, int32_t **adjwgt_out)
{
    if (!nBricks || !xadj_out || !adj_out || !vwgt_out || !adjwgt_out) return ehyb_fail(EHYB_ERR_ARG, "ehyb_grid_brick_graph: NULL");
    ehyb_grid_decomp *D = NULL;
    int rc = ehyb_grid_decomp_create(nx, ny, nz, bx, by, bz, 1, NULL, &D);
    if (rc) return rc;
    const int64_t nb = D->nb;
    uint32_t *xadj = (uint32_t *)malloc(((size_t)nb + 1) * sizeof(uint32_t));
    uint32_t *adj = (uint32_t *)malloc((size_t)nb * 26 * sizeof(uint32_t));
    int32_t *vw = (int32_t *)malloc((size_t)nb * sizeof(int32_t));
    int32_t *aw = (int32_t *)malloc((size_t)nb * 26 * sizeof(int32_t));
    if (!xadj || !adj || !vw || !aw) { free(xadj); free(adj); free(vw); free(aw); ehyb_grid_decomp_free(D); return ehyb_fail(EHYB_ERR_NOMEM, "grid: out of memory"); }
    uint32_t k = 0;
    for (int64_t b = 0; b < nb; ++b) {
        const brick_box e = brick_extent(D, b);
        const int ix = (int)(b % D->nbx), iy = (int)((b / D->nbx) % D->nby), iz = (int)(b / ((int64_t)D->nbx * D->nby));
        xadj[b] = k;
        vw[b] = e.sx * e.sy * e.sz;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy && !dz) continue;
                    const int jx = ix + dx, jy = iy + dy, jz = iz + dz;
                    if (jx < 0 || jx >= D->nbx || jy < 0 || jy >= D->nby || jz < 0 || jz >= D->nbz) continue;
                    adj[k] = (uint32_t)(((int64_t)jz * D->nby + jy) * D->nbx + jx);
                    aw[k] = (dx ? 1 : 3 * e.sx - 2) * (dy ? 1 : 3 * e.sy - 2) * (dz ? 1 : 3 * e.sz - 2);
                    ++k;
                }
    }
    xadj[nb] = k;
    ehyb_grid_decomp_free(D);
    *nBricks = nb; *xadj_out = xadj; *adj_out = adj; *vwgt_out = vw; *adjwgt_out = aw;
    return EHYB_OK;
}

/* number of stencil neighbours (self included) of coordinate v on an axis of length n */
static inline int axis_count(int v, int n) { return 1 + (v > 0) + (v < n - 1); }

/* entries of the row of cell (x, y, z): the 27-point stencil clipped at the domain boundary */
static inline int row_len(const ehyb_grid_decomp *D, int x, int y, int z)
{
    return axis_count(x, D->nx) * axis_count(y, D->ny) * axis_count(z, D->nz);
}

/* Sorts the cells of a brick by (entries inside the brick descending, natural index ascending):
 * perm[i] = position of natural in-brick cell i, inv[position] = cell.  The count of a cell is
 * the product of its per-axis counts inside the brick, at most 27: a counting sort. */
static void brick_perm(const brick_box *e, int32_t *perm, int32_t *inv)
{
    int hist[28];
    memset(hist, 0, sizeof hist);
    const int cells = e->sx * e->sy * e->sz;
    for (int i = 0; i < cells; ++i) {
        const int x = i % e->sx, y = (i / e->sx) % e->sy, z = i / (e->sx * e->sy);
        hist[axis_count(x, e->sx) * axis_count(y, e->sy) * axis_count(z, e->sz)] += 1;
    }
    int start[28], acc = 0;
    for (int c = 27; c >= 0; --c) { start[c] = acc; acc += hist[c]; }
    for (int i = 0; i < cells; ++i) {
        const int x = i % e->sx, y = (i / e->sx) % e->sy, z = i / (e->sx * e->sy);
        const int pos = start[axis_count(x, e->sx) * axis_count(y, e->sy) * axis_count(z, e->sz)]++;
        perm[i] = pos;
        if (inv) inv[pos] = i;
    }
}

typedef struct { int64_t *v; int64_t n, cap; } vec64;

static int vec64_push(vec64 *a, int64_t x)
{
    if (a->n == a->cap) {
        const int64_t nc = a->cap ? 2 * a->cap : 4096;
        int64_t *q = (int64_t *)realloc(a->v, (size_t)nc * sizeof(int64_t));
        if (!q) return -1;
        a->v = q; a->cap = nc;
    }
    a->v[a->n++] = x;
    return 0;
}

static int cmp_i64(const void *a, const void *b)
{
    const int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return x < y ? -1 : x > y;
}

/* All rows of a rank in level-1 order (cells of a brick in natural order), level-1 column ids,
 * entries of a row in ascending NATURAL column order ((dz, dy, dx) loops: the order of the
 * matrix in natural numbering): the input of the general path (ehyb_mg_local_build), for the
 * parity test of the streamed build and for small grids. */
int ehyb_grid_rows(const ehyb_grid_decomp *D, int rank, int64_t **rowPtr_out, int64_t **col_out, double **val_out, uint32_t **partVec_out)
{
    if (!D || rank < 0 || rank >= D->nranks || !rowPtr_out || !col_out || !val_out) return ehyb_fail(EHYB_ERR_ARG, "ehyb_grid_rows: bad argument");
    const int64_t n = D->rowStarts[rank + 1] - D->rowStarts[rank];
    if (n <= 0 || n > INT_MAX) return ehyb_fail(EHYB_ERR_ARG, "rank %d owns %lld rows", rank, (long long)n);
    int64_t *rowPtr = (int64_t *)malloc(((size_t)n + 1) * sizeof(int64_t));
    uint32_t *pv = partVec_out ? (uint32_t *)malloc((size_t)n * sizeof(uint32_t)) : NULL;
    if (!rowPtr || (partVec_out && !pv)) { free(rowPtr); free(pv); return ehyb_fail(EHYB_ERR_NOMEM, "grid: out of memory"); }
    rowPtr[0] = 0;
    int64_t r = 0;
    for (int64_t k = D->rankBrickStart[rank]; k < D->rankBrickStart[rank + 1]; ++k) {
        const brick_box e = brick_extent(D, D->rankBricks[k]);
        for (int z = e.z0; z < e.z0 + e.sz; ++z)
            for (int y = e.y0; y < e.y0 + e.sy; ++y)
                for (int x = e.x0; x < e.x0 + e.sx; ++x, ++r) {
                    rowPtr[r + 1] = rowPtr[r] + row_len(D, x, y, z);
                    if (pv) pv[r] = (uint32_t)(k - D->rankBrickStart[rank]);
                }
    }
    const int64_t nnz = rowPtr[n];
    int64_t *col = (int64_t *)malloc((size_t)nnz * sizeof(int64_t));
    double *val = (double *)malloc((size_t)nnz * sizeof(double));
    if (!col || !val) { free(rowPtr); free(pv); free(col); free(val); return ehyb_fail(EHYB_ERR_NOMEM, "grid: out of memory"); }
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t k = D->rankBrickStart[rank]; k < D->rankBrickStart[rank + 1]; ++k) {
        const int64_t b = D->rankBricks[k];
        const brick_box e = brick_extent(D, b);
        int64_t rr = D->base[b] - D->rowStarts[rank];
        for (int z = e.z0; z < e.z0 + e.sz; ++z)
            for (int y = e.y0; y < e.y0 + e.sy; ++y)
                for (int x = e.x0; x < e.x0 + e.sx; ++x, ++rr) {
                    int64_t o = rowPtr[rr];
                    for (int dz = -1; dz <= 1; ++dz)
                        for (int dy = -1; dy <= 1; ++dy)
                            for (int dx = -1; dx <= 1; ++dx) {
                                const int X = x + dx, Y = y + dy, Z = z + dz;
                                if (X < 0 || X >= D->nx || Y < 0 || Y >= D->ny || Z < 0 || Z >= D->nz) continue;
                                col[o] = level1_id(D, X, Y, Z);
                                val[o] = (dx | dy | dz) ? -1.0 : 26.0;
                                ++o;
                            }
                }
    }
    *rowPtr_out = rowPtr; *col_out = col; *val_out = val;
    if (partVec_out) *partVec_out = pv;
    return EHYB_OK;
}

/* natural grid index (z * ny + y) * nx + x of every local row in PERMUTED order (the order of the
 * block's x and y vectors): what a caller needs to fill x and to check y */
int ehyb_mg_local_natural_ids(const ehyb_mg_local *L, int64_t *out)
{
    if (!L || !out || !L->streamed || !L->grid) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_local_natural_ids: not a grid block");
    const ehyb_grid_decomp *D = L->grid;
    const int rank = L->rank;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t k = D->rankBrickStart[rank]; k < D->rankBrickStart[rank + 1]; ++k) {
        const int64_t b = D->rankBricks[k];
        const brick_box e = brick_extent(D, b);
        int64_t l = D->base[b] - D->rowStarts[rank];
        for (int z = e.z0; z < e.z0 + e.sz; ++z)
            for (int y = e.y0; y < e.y0 + e.sy; ++y)
                for (int x = e.x0; x < e.x0 + e.sx; ++x, ++l) out[L->perm[l]] = ((int64_t)z * D->ny + y) * D->nx + x;
    }
    return EHYB_OK;
}

/* natural grid index of level-1 ids (the halo list of a block) */
int ehyb_grid_natural_ids(const ehyb_grid_decomp *D, int64_t count, const int64_t *level1, int64_t *out)
{
    if (!D || count < 0 || (count && (!level1 || !out))) return ehyb_fail(EHYB_ERR_ARG, "ehyb_grid_natural_ids: bad argument");
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t i = 0; i < count; ++i) {
        const int64_t id = level1[i];
        if (id < 0 || id >= D->rowStarts[D->nranks]) { bad = 1; continue; }
        /* owner by the rank ranges, then its brick by bisection over the rank's ascending bases */
        int g = 0;
        while (id >= D->rowStarts[g + 1]) ++g;
        int64_t lo = D->rankBrickStart[g], hi = D->rankBrickStart[g + 1] - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) / 2;
            if (D->base[D->rankBricks[mid]] <= id) lo = mid; else hi = mid - 1;
        }
        const int64_t b = D->rankBricks[lo];
        const brick_box e = brick_extent(D, b);
        const int64_t off = id - D->base[b];
        const int x = e.x0 + (int)(off % e.sx), y = e.y0 + (int)((off / e.sx) % e.sy), z = e.z0 + (int)(off / ((int64_t)e.sx * e.sy));
        out[i] = ((int64_t)z * D->ny + y) * D->nx + x;
    }
    return bad ? ehyb_fail(EHYB_ERR_ARG, "ehyb_grid_natural_ids: id outside the grid") : EHYB_OK;
}

/*
 * A rank's block, streamed (see the head of the file).  The partitions are the rank's bricks
 * (window W = cells of a full brick rounded up to 64), exchange as in ehyb_mg_local_finish,
 * chunkBricks bricks per call of the layout builder (0 = 64).  The block comes out finished
 * (layout built, permutation known); the caller still exchanges the halo lists with the other
 * ranks (ehyb_mg_local_halo -> ehyb_mg_local_set_send) before it creates the device session.
 * D must outlive the block.
 */
int ehyb_mg_grid_build(const ehyb_grid_decomp *D, int rank, double er_fill, int exchange, int chunkBricks, ehyb_mg_local **out)
{
    if (!D || !out || rank < 0 || rank >= D->nranks || (exchange != EHYB_MG_NCCL && exchange != EHYB_MG_P2P))
        return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_grid_build: bad argument");
    const int64_t r0 = D->rowStarts[rank], n64 = D->rowStarts[rank + 1] - r0;
    const int64_t k0 = D->rankBrickStart[rank], K = D->rankBrickStart[rank + 1] - k0;
    if (n64 <= 0 || K <= 0) return ehyb_fail(EHYB_ERR_ARG, "rank %d owns no brick", rank);
    if (n64 > INT_MAX) return ehyb_fail(EHYB_ERR_LIMIT, "rank %d owns %lld rows: row indices are 32-bit per GPU", rank, (long long)n64);
    if (chunkBricks <= 0) chunkBricks = 64;
    const int n = (int)n64;
    const int W = (int)ehyb_round_up64((int64_t)D->bx * D->by * D->bz, 64);
    int rc = EHYB_OK;
    ehyb_mg_local *L = (ehyb_mg_local *)calloc(1, sizeof *L);
    if (!L) return ehyb_fail(EHYB_ERR_NOMEM, "grid: out of memory");
    L->rank = rank; L->nranks = D->nranks; L->n = n; L->streamed = 1; L->finished = 1; L->grid = D;
    L->rowStarts = (int64_t *)malloc(((size_t)D->nranks + 1) * sizeof(int64_t));
    L->recvCount = (int64_t *)calloc((size_t)D->nranks, sizeof(int64_t));
    L->sendCount = (int64_t *)calloc((size_t)D->nranks, sizeof(int64_t));
    L->perm = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    vec64 *ext = (vec64 *)calloc((size_t)nthreads, sizeof(vec64));
    int64_t *rowPtr = NULL;
    int32_t *col = NULL, *pb = NULL;
    double *val = NULL;
    ehyb_layout_builder *B = NULL;
    if (!L->rowStarts || !L->recvCount || !L->sendCount || !L->perm || !ext) { rc = ehyb_fail(EHYB_ERR_NOMEM, "grid: out of memory"); goto fail; }
    memcpy(L->rowStarts, D->rowStarts, ((size_t)D->nranks + 1) * sizeof(int64_t));

    /* ---- pass A: permutation inside every brick; level-1 ids referenced on other ranks ---- */
    int oom = 0;
    int64_t nnz = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(| : oom) reduction(+ : nnz)
    for (int64_t k = 0; k < K; ++k) {
        const int64_t b = D->rankBricks[k0 + k];
        const brick_box e = brick_extent(D, b);
        const int64_t lb = D->base[b] - r0;
        brick_perm(&e, L->perm + lb, NULL);
        const int cells = e.sx * e.sy * e.sz;
        for (int i = 0; i < cells; ++i) L->perm[lb + i] += (int32_t)lb;
        /* entries of the brick's rows (closed form per axis) */
        int64_t ax = 0, ay = 0, az = 0;
        for (int x = e.x0; x < e.x0 + e.sx; ++x) ax += axis_count(x, D->nx);
        for (int y = e.y0; y < e.y0 + e.sy; ++y) ay += axis_count(y, D->ny);
        for (int z = e.z0; z < e.z0 + e.sz; ++z) az += axis_count(z, D->nz);
        nnz += ax * ay * az;
        /* does any of the 26 neighbour bricks belong to another rank? */
        const int ix = (int)(b % D->nbx), iy = (int)((b / D->nbx) % D->nby), iz = (int)(b / ((int64_t)D->nbx * D->nby));
        int foreign = 0;
        for (int dz = -1; dz <= 1 && !foreign; ++dz)
            for (int dy = -1; dy <= 1 && !foreign; ++dy)
                for (int dx = -1; dx <= 1 && !foreign; ++dx) {
                    const int jx = ix + dx, jy = iy + dy, jz = iz + dz;
                    if (jx < 0 || jx >= D->nbx || jy < 0 || jy >= D->nby || jz < 0 || jz >= D->nbz) continue;
                    foreign = D->owner[((int64_t)jz * D->nby + jy) * D->nbx + jx] != rank;
                }
        if (!foreign) continue;
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        for (int z = e.z0; z < e.z0 + e.sz; ++z)
            for (int y = e.y0; y < e.y0 + e.sy; ++y)
                for (int x = e.x0; x < e.x0 + e.sx; ++x) {
                    if (x > e.x0 && x < e.x0 + e.sx - 1 && y > e.y0 && y < e.y0 + e.sy - 1 && z > e.z0 && z < e.z0 + e.sz - 1) continue;
                    for (int dz = -1; dz <= 1; ++dz)
                        for (int dy = -1; dy <= 1; ++dy)
                            for (int dx = -1; dx <= 1; ++dx) {
                                const int X = x + dx, Y = y + dy, Z = z + dz;
                                if (X < 0 || X >= D->nx || Y < 0 || Y >= D->ny || Z < 0 || Z >= D->nz) continue;
                                if (D->owner[brick_of_cell(D, X, Y, Z)] != rank) oom |= vec64_push(&ext[t], level1_id(D, X, Y, Z)) != 0;
                            }
                }
    }
    if (oom) { rc = ehyb_fail(EHYB_ERR_NOMEM, "grid: out of memory"); goto fail; }
    L->nnz = nnz;
    {   /* halo = sorted unique ids; sorted level-1 ids come grouped by owner */
        int64_t tot = 0;
        for (int t = 0; t < nthreads; ++t) tot += ext[t].n;
        int64_t *all = (int64_t *)malloc((size_t)(tot ? tot : 1) * sizeof(int64_t));
        if (!all) { rc = ehyb_fail(EHYB_ERR_NOMEM, "grid: out of memory"); goto fail; }
        int64_t o = 0;
        for (int t = 0; t < nthreads; ++t) {
            if (ext[t].n) memcpy(all + o, ext[t].v, (size_t)ext[t].n * sizeof(int64_t));
            o += ext[t].n;
            free(ext[t].v);
            ext[t].v = NULL;
        }
        qsort(all, (size_t)tot, sizeof(int64_t), cmp_i64);
        int64_t nHalo = 0;
        for (int64_t i = 0; i < tot; ++i)
            if (i == 0 || all[i] != all[i - 1]) all[nHalo++] = all[i];
        L->nHalo = nHalo;
        L->haloGlobal = all;
        int owner = 0;
        for (int64_t i = 0; i < nHalo; ++i) {
            while (all[i] >= D->rowStarts[owner + 1]) ++owner;
            L->recvCount[owner] += 1;
        }
        if ((int64_t)n + nHalo > INT_MAX) { rc = ehyb_fail(EHYB_ERR_LIMIT, "local block + halo exceed 2^31 columns"); goto fail; }
    }

    /* ---- pass B: permuted rows, chunkBricks bricks at a time, into the layout builder ---- */
    {
        ehyb_layout_opts o;
        memset(&o, 0, sizeof o);
        o.W = W; o.ctasPerPart = 1; o.er_fill = er_fill; o.ncols = (int64_t)n + L->nHalo;
        o.halo_in_overflow = exchange == EHYB_MG_NCCL;
        rc = ehyb_layout_builder_begin(n, &o, &B);
        if (rc) goto fail;
    }
    const int64_t chunkRowsMax = (int64_t)chunkBricks * D->bx * D->by * D->bz;
    rowPtr = (int64_t *)malloc(((size_t)chunkRowsMax + 1) * sizeof(int64_t));
    col = (int32_t *)malloc((size_t)chunkRowsMax * 27 * sizeof(int32_t));
    val = (double *)malloc((size_t)chunkRowsMax * 27 * sizeof(double));
    pb = (int32_t *)malloc(((size_t)chunkBricks + 1) * sizeof(int32_t));
    if (!rowPtr || !col || !val || !pb) { rc = ehyb_fail(EHYB_ERR_NOMEM, "grid: out of memory"); goto fail; }
    for (int64_t kc = 0; kc < K && rc == EHYB_OK; kc += chunkBricks) {
        const int nk = (int)(K - kc < chunkBricks ? K - kc : chunkBricks);
        const int64_t chunkBase = D->base[D->rankBricks[k0 + kc]] - r0; /* bricks of a rank are consecutive in level-1 order */
        /* row lengths in permuted order, brick by brick */
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = 0; k < nk; ++k) {
            const int64_t b = D->rankBricks[k0 + kc + k];
            const brick_box e = brick_extent(D, b);
            const int64_t lb = D->base[b] - r0;
            pb[k] = (int32_t)lb;
            const int cells = e.sx * e.sy * e.sz;
            for (int i = 0; i < cells; ++i) {
                const int x = e.x0 + i % e.sx, y = e.y0 + (i / e.sx) % e.sy, z = e.z0 + i / (e.sx * e.sy);
                rowPtr[L->perm[lb + i] - chunkBase + 1] = row_len(D, x, y, z);
            }
        }
        const int64_t lastB = D->rankBricks[k0 + kc + nk - 1];
        const brick_box le = brick_extent(D, lastB);
        const int64_t chunkRows = D->base[lastB] - r0 + (int64_t)le.sx * le.sy * le.sz - chunkBase;
        pb[nk] = (int32_t)(chunkBase + chunkRows);
        rowPtr[0] = 0;
        for (int64_t r = 0; r < chunkRows; ++r) rowPtr[r + 1] += rowPtr[r];
        /* the entries */
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = 0; k < nk; ++k) {
            const int64_t b = D->rankBricks[k0 + kc + k];
            const brick_box e = brick_extent(D, b);
            const int64_t lb = D->base[b] - r0;
            const int cells = e.sx * e.sy * e.sz;
            const int32_t *pm = L->perm + lb; /* in-brick cell -> permuted local row */
            for (int i = 0; i < cells; ++i) {
                const int cx = i % e.sx, cy = (i / e.sx) % e.sy, cz = i / (e.sx * e.sy);
                const int x = e.x0 + cx, y = e.y0 + cy, z = e.z0 + cz;
                int64_t o = rowPtr[pm[i] - chunkBase];
                if (cx > 0 && cx < e.sx - 1 && cy > 0 && cy < e.sy - 1 && cz > 0 && cz < e.sz - 1) {
                    /* interior cell of the brick: all 27 neighbours are cells of this brick */
                    for (int dz = -1; dz <= 1; ++dz)
                        for (int dy = -1; dy <= 1; ++dy) {
                            const int j = i + (dz * e.sy + dy) * e.sx;
                            col[o] = pm[j - 1]; col[o + 1] = pm[j]; col[o + 2] = pm[j + 1];
                            val[o] = -1.0; val[o + 1] = (dz | dy) ? -1.0 : 26.0; val[o + 2] = -1.0;
                            o += 3;
                        }
                    continue;
                }
                for (int dz = -1; dz <= 1; ++dz)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int X = x + dx, Y = y + dy, Z = z + dz;
                            if (X < 0 || X >= D->nx || Y < 0 || Y >= D->ny || Z < 0 || Z >= D->nz) continue;
                            const int64_t id = level1_id(D, X, Y, Z);
                            if (id >= r0 && id < r0 + n) {
                                col[o] = L->perm[id - r0];
                            } else {
                                int64_t lo = 0, hi = L->nHalo - 1;
                                while (lo < hi) {
                                    const int64_t mid = (lo + hi) / 2;
                                    if (L->haloGlobal[mid] < id) lo = mid + 1; else hi = mid;
                                }
                                col[o] = (int32_t)(n + lo);
                            }
                            val[o] = (dx | dy | dz) ? -1.0 : 26.0;
                            ++o;
                        }
            }
        }
        rc = ehyb_layout_builder_add(B, nk, pb, rowPtr, col, val);
    }
    if (rc) goto fail;
    rc = ehyb_layout_builder_finish(B, &L->layout);
    B = NULL;
    if (rc) goto fail;
    free(rowPtr); free(col); free(val); free(pb); free(ext);
    *out = L;
    return EHYB_OK;

fail:
    if (ext) for (int t = 0; t < nthreads; ++t) free(ext[t].v);
    free(ext); free(rowPtr); free(col); free(val); free(pb);
    ehyb_layout_builder_abort(B);
    ehyb_mg_local_free(L);
    return rc;
}
