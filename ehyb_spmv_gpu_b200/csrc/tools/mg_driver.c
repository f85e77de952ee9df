/*
 * mg_driver.c -- the multi-GPU mode of the driver:  ./spmv.out -G <gpus> -g st27:NX:NY:NZ -i <iterations>
 *
 * BASELINE.json config 5 from C, with no launcher, no Python and no torch.distributed: ONE process,
 * one host thread per GPU.  The 27-point stencil grid is cut into bricks (16^3 cells unless -B says
 * otherwise); level 1 = the pinned mt-metis k = G partition of the weighted brick graph (bricks ->
 * GPUs; the call of reordering.c:270-293 on the coarsened graph), level 2 = one EHYB partition per
 * brick; every thread streams its GPU's block into the tuned layout (csrc/host/grid.c), the halo
 * lists are exchanged through memory, the sessions reach each other's halo buffers by plain peer
 * access (ehyb_mg_p2p_connect_local) and exchange x inside the persistent kernel every product.
 *
 * Like the reference's driver (solver_test.c:228-232, :247) it makes its own x and its own check
 * vector on the CPU - for a matrix that never exists as a whole both are functions of the grid
 * index: x_i = hash(i) in (-0.1, 0.1), y_i = 26 x_i - sum of the neighbours - and compares every row
 * of every GPU against the accuracy gate |y - y_ref| <= 1e-12 (|A||x|); the exit code tells.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>
#include <omp.h>
#include "ehyb.h"

static double now_s(void)
{
    struct timeval t;
    gettimeofday(&t, NULL);
    return (double)t.tv_sec + 1e-6 * (double)t.tv_usec;
}

/* x as a function of the natural grid index (splitmix-style hash, in (-0.1, 0.1)) */
static inline double x_of(uint64_t i)
{
    uint64_t z = i * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL;
    z ^= z >> 29;
    z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 32;
    return ((double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 0.2;
}

typedef struct {
    ehyb_mg_local *loc;
    ehyb_mg_session *ses;
    int64_t n, nnz, nHalo, algBytes;
    const int64_t *halo, *recvCount;
    int peers, gateFail, rc;
    float ms;
    double buildS;
    char err[512];
    char kernel[64];
} rank_state;

int ehyb_driver_multi_gpu(int G, const char *gen, const char *brickSpec, int iters)
{
    char kind[32] = "", buf[256];
    int nx = 0, ny = 0, nz = 0, bx = 16, by = 16, bz = 16;
    snprintf(buf, sizeof buf, "%s", gen);
    for (char *p = buf; *p; ++p) if (*p == ':') *p = ' ';
    if (sscanf(buf, "%31s %d %d %d", kind, &nx, &ny, &nz) != 4 || strcmp(kind, "st27") != 0 || nx <= 0 || ny <= 0 || nz <= 0) {
        printf("-G needs -g st27:NX:NY:NZ (the sharded 27-point stencil, BASELINE.json config 5)\n");
        return 1;
    }
    if (brickSpec && brickSpec[0] && sscanf(brickSpec, "%dx%dx%d", &bx, &by, &bz) != 3) { printf("-B wants BXxBYxBZ\n"); return 1; }
    int devs = 0;
    if (ehyb_device_count(&devs) || devs < G) { printf("%d GPUs asked for, %d present\n", G, devs); return 1; }
    for (int g = 0; g < G; ++g) {
        int ok = 0;
        if (ehyb_mg_p2p_supported(g, G, &ok) || !ok) { printf("GPU %d has no peer access to the others: the in-kernel exchange needs NVLink / NVSwitch\n", g); return 1; }
    }
    const int64_t N = (int64_t)nx * ny * nz;
    printf("multi-GPU: 27-point stencil %d x %d x %d (n %lld), %d GPUs, bricks %d x %d x %d\n", nx, ny, nz, (long long)N, G, bx, by, bz);

    /* ---- level 1: bricks -> GPUs ---- */
    const double t0 = now_s();
    int64_t nb = 0;
    uint32_t *xadj = NULL, *adj = NULL, *owner = NULL;
    int32_t *vw = NULL, *aw = NULL;
    if (ehyb_grid_brick_graph(nx, ny, nz, bx, by, bz, &nb, &xadj, &adj, &vw, &aw)) { printf("brick graph: %s\n", ehyb_last_error()); return 1; }
    owner = (uint32_t *)calloc((size_t)nb, sizeof(uint32_t));
    printf("start k-way partition\n");
    if (G > 1 && ehyb_partition_graph_weighted((uint32_t)nb, xadj, adj, vw, aw, (uint32_t)G, 1, 1.001f, owner)) { printf("level-1 partition: %s\n", ehyb_last_error()); return 1; }
    printf("partition finished\n");
    ehyb_free_host(xadj); ehyb_free_host(adj); ehyb_free_host(vw); ehyb_free_host(aw);
    ehyb_grid_decomp *D = NULL;
    if (ehyb_grid_decomp_create(nx, ny, nz, bx, by, bz, G, G > 1 ? owner : NULL, &D)) { printf("decomposition: %s\n", ehyb_last_error()); return 1; }
    free(owner);
    printf("level 1: %lld bricks on %d GPUs by mt-metis (%.2f s)\n", (long long)nb, G, now_s() - t0);

    rank_state *R = (rank_state *)calloc((size_t)G, sizeof(rank_state));
    ehyb_mg_session **S = (ehyb_mg_session **)calloc((size_t)G, sizeof(ehyb_mg_session *));
    const int cores = omp_get_num_procs();
    const int inner = cores / G > 0 ? cores / G : 1;
    omp_set_max_active_levels(2);
    int failed = 0;

    /* ---- every GPU's block, streamed (level 2 + format build), in parallel ---- */
#pragma omp parallel num_threads(G)
    {
        const int r = omp_get_thread_num();
        rank_state *q = &R[r];
        ehyb_set_host_threads(inner);
        const double tb = now_s();
        q->rc = ehyb_mg_grid_build(D, r, 0.0, EHYB_MG_P2P, 0, &q->loc);
        if (q->rc) snprintf(q->err, sizeof q->err, "%s", ehyb_last_error());
        else ehyb_mg_local_halo(q->loc, &q->nHalo, &q->halo, &q->recvCount);
        q->buildS = now_s() - tb;
    }
    for (int r = 0; r < G; ++r) if (R[r].rc) { printf("GPU %d: format build: %s\n", r, R[r].err); failed = 1; }
    if (failed) return 1;

    /* ---- who sends what: rank g sends to rank r the part of r's halo list that g owns ---- */
    for (int g = 0; g < G; ++g) {
        int64_t *cnt = (int64_t *)calloc((size_t)G, sizeof(int64_t)), tot = 0;
        for (int r = 0; r < G; ++r) { cnt[r] = R[r].recvCount[g]; tot += cnt[r]; }
        int64_t *ids = (int64_t *)malloc((size_t)(tot ? tot : 1) * sizeof(int64_t)), o = 0;
        for (int r = 0; r < G; ++r) {
            int64_t off = 0;
            for (int k = 0; k < g; ++k) off += R[r].recvCount[k];
            memcpy(ids + o, R[r].halo + off, (size_t)cnt[r] * sizeof(int64_t));
            o += cnt[r];
        }
        if (ehyb_mg_local_set_send(R[g].loc, cnt, ids)) { printf("GPU %d: send list: %s\n", g, ehyb_last_error()); return 1; }
        free(cnt); free(ids);
    }

    /* ---- sessions (one per GPU, created by its thread), peer connection ---- */
#pragma omp parallel num_threads(G)
    {
        const int r = omp_get_thread_num();
        R[r].rc = ehyb_mg_session_create_p2p(R[r].loc, r, G, r, &S[r]);
        if (R[r].rc) snprintf(R[r].err, sizeof R[r].err, "%s", ehyb_last_error());
    }
    for (int r = 0; r < G; ++r) if (R[r].rc) { printf("GPU %d: session: %s\n", r, R[r].err); failed = 1; }
    if (failed) return 1;
    if (ehyb_mg_p2p_connect_local(S, G)) { printf("peer connection: %s\n", ehyb_last_error()); return 1; }
    const double tPrep = now_s() - t0;

    /* ---- x, one checked product, the timed loop: every GPU from its own thread ---- */
#pragma omp parallel num_threads(G)
    {
        const int r = omp_get_thread_num();
        rank_state *q = &R[r];
        ehyb_set_host_threads(inner);
        ehyb_handle *h = NULL;
        ehyb_mg_session_handle(S[r], &h);
        int64_t n = 0, ncols = 0;
        ehyb_session_size(h, &n, &ncols);
        q->n = n;
        const ehyb_layout *lay = NULL;
        ehyb_mg_local_view(q->loc, NULL, &lay, NULL, NULL, NULL);
        ehyb_layout_view v;
        ehyb_layout_get(lay, &v);
        q->nnz = v.nnz; q->algBytes = v.algBytes;
        for (int g = 0; g < G; ++g) q->peers += q->recvCount[g] > 0;
        snprintf(q->kernel, sizeof q->kernel, "%s", ehyb_session_kernel(h));
        int64_t *nat = (int64_t *)malloc((size_t)n * sizeof(int64_t));
        double *x = (double *)calloc((size_t)ncols, sizeof(double)), *y = (double *)malloc((size_t)n * sizeof(double));
        double *xd = NULL, *yd = NULL;
        q->rc = (!nat || !x || !y) ? EHYB_ERR_NOMEM : ehyb_mg_local_natural_ids(q->loc, nat);
        if (!q->rc) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) x[i] = x_of((uint64_t)nat[i]);
            q->rc = ehyb_set_x(h, x);
        }
        if (!q->rc) q->rc = ehyb_session_vectors(h, &xd, &yd);
#pragma omp barrier
        if (!q->rc) q->rc = ehyb_mg_spmv(S[r], xd, yd);
        if (!q->rc) q->rc = ehyb_get_y(h, y);
        if (!q->rc) {
            /* the driver's own check vector, in closed form from the generator's definition */
            const int64_t plane = (int64_t)nx * ny;
            int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
            for (int64_t i = 0; i < n; ++i) {
                const int64_t g = nat[i];
                const int cx = (int)(g % nx), cy = (int)((g / nx) % ny), cz = (int)(g / plane);
                double s = 0.0, a = 0.0;
                for (int dz = -1; dz <= 1; ++dz)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            if (cx + dx < 0 || cx + dx >= nx || cy + dy < 0 || cy + dy >= ny || cz + dz < 0 || cz + dz >= nz) continue;
                            const int64_t c = g + dz * plane + (int64_t)dy * nx + dx;
                            const double val = c == g ? 26.0 : -1.0, xv = x_of((uint64_t)c);
                            s += val * xv;
                            a += fabs(val) * fabs(xv);
                        }
                bad += !(fabs(y[i] - s) <= 1e-12 * a);
            }
            q->gateFail = bad;
        }
#pragma omp barrier
        if (!q->rc) q->rc = ehyb_mg_time_spmv(S[r], 10, iters, &q->ms);
        if (q->rc) snprintf(q->err, sizeof q->err, "%s", ehyb_last_error());
        free(nat); free(x); free(y);
    }

    /* ---- report ---- */
    double msMax = 0.0;
    int64_t nnzAll = 0, gate = 0, haloAll = 0;
    for (int r = 0; r < G; ++r) {
        if (R[r].rc) { printf("GPU %d: %s\n", r, R[r].err); failed = 1; continue; }
        const double us = R[r].ms / iters * 1e3;
        printf("GPU %d: %lld rows, %lld entries, halo %lld x entries from %d peers, build %.1f s, kernel %s: %.1f us per product, %.1f GFLOP/s, "
               "%.1f GB/s algorithmic, gate %d rows fail\n", r, (long long)R[r].n, (long long)R[r].nnz, (long long)R[r].nHalo, R[r].peers, R[r].buildS,
               R[r].kernel, us, 2.0 * (double)R[r].nnz / (us * 1e3), (double)R[r].algBytes / (us * 1e3), R[r].gateFail);
        if (R[r].ms > msMax) msMax = R[r].ms;
        nnzAll += R[r].nnz; gate += R[r].gateFail; haloAll += R[r].nHalo;
    }
    if (!failed) {
        const double us = msMax / iters * 1e3;
        printf("iter is %d, time is %f ms, GPU Gflops is %f\n", iters, msMax, 2.0 * (double)nnzAll / (us * 1e3));
        printf("EHYB-B200 multi-GPU: %d GPUs, %lld entries, %lld halo x entries per product, %.1f us per product (slowest GPU), %.1f GFLOP/s, host set-up %.1f s\n",
               G, (long long)nnzAll, (long long)haloAll, us, 2.0 * (double)nnzAll / (us * 1e3), tPrep);
        printf("accuracy gate |y - y_ref| <= 1e-12*(|A||x|): %lld of %lld rows fail\n", (long long)gate, (long long)N);
    }
    /* every rank has finished its products (the parallel region ended): free */
    for (int r = 0; r < G; ++r) { ehyb_mg_session_free(S[r]); ehyb_mg_local_free(R[r].loc); }
    ehyb_grid_decomp_free(D);
    free(R); free(S);
    return failed ? 1 : (gate ? 3 : 0);
}
