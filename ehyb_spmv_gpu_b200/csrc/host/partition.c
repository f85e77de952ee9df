/*
 * partition.c -- partitioner hook.
 *
 * The partition vector comes from mt-metis, a pinned third-party binary (see
 * third_party/mtmetis/mtmetis_abi.h).  Because that archive is not position independent the
 * shared library reaches it out of process: the graph goes to a temporary file,
 * bin/ehyb_mtmetis (tools/mtmetis_helper.c) makes the reference's MTMETIS_PartGraphKway call
 * (reordering.c:270-293) and the partition vector comes back through a second file.
 * Executables that link libmtmetis.a (bin/spmv.out) install a direct call with
 * ehyb_set_partitioner().
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/wait.h>
#include "common.h"

static ehyb_partition_fn g_fn = NULL;
static void *g_user = NULL;

void ehyb_set_partitioner(ehyb_partition_fn fn, void *user)
{
    g_fn = fn;
    g_user = user;
}

static int find_helper(char *out, size_t cap)
{
    const char *env = getenv("EHYB_MTMETIS_BIN");
    if (env && env[0]) {
        snprintf(out, cap, "%s", env);
        return access(out, X_OK) == 0 ? 0 : -1;
    }
    /* <repo>/ehyb_spmv_gpu_b200/lib/libehyb.so -> <repo>/bin/ehyb_mtmetis */
    Dl_info info;
    if (dladdr((void *)&find_helper, &info) && info.dli_fname) {
        char dir[900];
        snprintf(dir, sizeof dir, "%s", info.dli_fname);
        for (int up = 0; up < 3; ++up) {
            char *s = strrchr(dir, '/');
            if (!s) break;
            *s = 0;
            snprintf(out, cap, "%s/bin/ehyb_mtmetis", dir);
            if (access(out, X_OK) == 0) return 0;
        }
    }
    return -1;
}

/* vwgt / adjwgt: both NULL (the reference's call) or both given (coarsened graphs, grid.c) */
static int helper_partition(uint32_t n, const uint32_t *xadj, const uint32_t *adjncy, const int32_t *vwgt, const int32_t *adjwgt,
                            uint32_t nparts, uint32_t nthreads, float ub, uint32_t *where)
{
    char bin[1024];
    if (find_helper(bin, sizeof bin))
        return ehyb_fail(EHYB_ERR_PARTITION,
                         "mt-metis helper not found (set EHYB_MTMETIS_BIN or build bin/ehyb_mtmetis)");
    char gpath[] = "/tmp/ehyb_graph_XXXXXX", wpath[] = "/tmp/ehyb_where_XXXXXX";
    int gfd = mkstemp(gpath), wfd = mkstemp(wpath);
    if (gfd < 0 || wfd < 0) return ehyb_fail(EHYB_ERR_IO, "mkstemp failed");
    close(wfd);
    FILE *f = fdopen(gfd, "wb");
    uint32_t hdr[4] = {vwgt ? 0x57594845u /* 'EHYW': weights follow */ : 0x47594845u /* 'EHYG' */, n, nparts, nthreads};
    int bad = fwrite(hdr, sizeof hdr, 1, f) != 1 || fwrite(&ub, 4, 1, f) != 1 ||
              fwrite(xadj, 4, (size_t)n + 1, f) != (size_t)n + 1 ||
              fwrite(adjncy, 4, xadj[n], f) != xadj[n];
    if (!bad && vwgt) bad = fwrite(vwgt, 4, n, f) != n || fwrite(adjwgt, 4, xadj[n], f) != xadj[n];
    bad |= fclose(f) != 0;
    int rc = EHYB_OK;
    if (bad) {
        rc = ehyb_fail(EHYB_ERR_IO, "writing %s failed", gpath);
    } else {
        pid_t pid = fork();
        if (pid == 0) {
            execl(bin, bin, gpath, wpath, (char *)NULL);
            _exit(127);
        }
        int st = 0;
        if (pid < 0 || waitpid(pid, &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) {
            rc = ehyb_fail(EHYB_ERR_PARTITION, "%s failed (status %d)", bin, st);
        } else {
            f = fopen(wpath, "rb");
            int32_t cut;
            if (!f || fread(&cut, 4, 1, f) != 1 || fread(where, 4, n, f) != n)
                rc = ehyb_fail(EHYB_ERR_IO, "reading %s failed", wpath);
            if (f) fclose(f);
        }
    }
    unlink(gpath);
    unlink(wpath);
    return rc;
}

int ehyb_partition_graph(uint32_t n, const uint32_t *xadj, const uint32_t *adjncy, uint32_t nparts,
                         uint32_t nthreads, uint32_t *where)
{
    if (!xadj || !adjncy || !where || nparts == 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_partition_graph: bad argument");
    if (nparts == 1) { /* nothing to partition (k-way partitioners reject or mishandle k = 1) */
        memset(where, 0, (size_t)n * sizeof(uint32_t));
        return EHYB_OK;
    }
    const float ub = 1.001f; /* reordering.c:273 */
    int rc = g_fn ? g_fn(n, xadj, adjncy, nparts, nthreads, ub, where, g_user)
                  : helper_partition(n, xadj, adjncy, NULL, NULL, nparts, nthreads, ub, where);
    if (rc) return rc;
    for (uint32_t i = 0; i < n; ++i)
        if (where[i] >= nparts) return ehyb_fail(EHYB_ERR_PARTITION, "partitioner returned part %u >= %u", where[i], nparts);
    return EHYB_OK;
}

/* The unweighted call through the helper process even when an in-process partitioner is installed:
 * hierpart.c runs several of these at the same time, and the in-process library call is not
 * re-entrant. */
int ehyb_partition_graph_process(uint32_t n, const uint32_t *xadj, const uint32_t *adjncy, uint32_t nparts, uint32_t *where)
{
    if (!xadj || !adjncy || !where || nparts == 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_partition_graph: bad argument");
    if (nparts == 1) {
        memset(where, 0, (size_t)n * sizeof(uint32_t));
        return EHYB_OK;
    }
    int rc = helper_partition(n, xadj, adjncy, NULL, NULL, nparts, 1, 1.001f, where);
    if (rc) return rc;
    for (uint32_t i = 0; i < n; ++i)
        if (where[i] >= nparts) return ehyb_fail(EHYB_ERR_PARTITION, "partitioner returned part %u >= %u", where[i], nparts);
    return EHYB_OK;
}

/* The same call with vertex and edge weights: the level-1 partition of a COARSENED graph (bricks
 * of a grid, grid.c) into one block per GPU.  Always through the helper binary. */
int ehyb_partition_graph_weighted(uint32_t n, const uint32_t *xadj, const uint32_t *adjncy, const int32_t *vwgt, const int32_t *adjwgt,
                                  uint32_t nparts, uint32_t nthreads, float ubvec, uint32_t *where)
{
    if (!xadj || !adjncy || !vwgt || !adjwgt || !where || nparts == 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_partition_graph_weighted: bad argument");
    if (nparts == 1) {
        memset(where, 0, (size_t)n * sizeof(uint32_t));
        return EHYB_OK;
    }
    int rc = helper_partition(n, xadj, adjncy, vwgt, adjwgt, nparts, nthreads, ubvec > 1.0f ? ubvec : 1.001f, where);
    if (rc) return rc;
    for (uint32_t i = 0; i < n; ++i)
        if (where[i] >= nparts) return ehyb_fail(EHYB_ERR_PARTITION, "partitioner returned part %u >= %u", where[i], nparts);
    return EHYB_OK;
}

int ehyb_partition_blocks(uint32_t n, uint32_t nparts, uint32_t *where)
{
    if (!where || nparts == 0) return ehyb_fail(EHYB_ERR_ARG, "ehyb_partition_blocks: bad argument");
    for (uint32_t p = 0; p < nparts; ++p) {
        uint64_t a = (uint64_t)n * p / nparts, b = (uint64_t)n * (p + 1) / nparts;
        for (uint64_t i = a; i < b; ++i) where[i] = p;
    }
    return EHYB_OK;
}
