/*
 * mtmetis_helper.c -- out-of-process front end to the pinned mt-metis binary.
 *
 * libmtmetis.a is not position independent, so libehyb.so cannot contain it.  The
 * library's default partitioner (host/partition.c) writes the graph to a file, runs this
 * program and reads the partition vector back; bin/spmv.out links mt-metis directly.
 * The call is the reference's: reordering.c:270-293 (symmetric, 1 thread) and :116-139
 * (unsymmetric, 6 threads): ncon 1, no weights, ubvec 1.001, default options otherwise.
 *
 * usage: ehyb_mtmetis <graph.bin> <where.bin>
 *   graph.bin : u32 magic 'EHYG' | 'EHYW', u32 nvtxs, u32 nparts, u32 nthreads, f32 ubvec,
 *               u32 xadj[nvtxs+1], u32 adjncy[xadj[nvtxs]]
 *               'EHYW' only (coarsened graphs: the level-1 partition into GPU blocks):
 *               i32 vwgt[nvtxs], i32 adjwgt[xadj[nvtxs]]
 *   where.bin : i32 edgecut, u32 where[nvtxs]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mtmetis_abi.h"

#define EHYG_MAGIC 0x47594845u
#define EHYW_MAGIC 0x57594845u

static int read_all(FILE *f, void *p, size_t bytes)
{
    return fread(p, 1, bytes, f) == bytes ? 0 : -1;
}

int main(int argc, char **argv)
{
    if (argc != 3) {
        fprintf(stderr, "usage: %s graph.bin where.bin\n", argv[0]);
        return 2;
    }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    uint32_t hdr[4];
    float ub;
    if (read_all(f, hdr, sizeof hdr) || read_all(f, &ub, sizeof ub) || (hdr[0] != EHYG_MAGIC && hdr[0] != EHYW_MAGIC)) {
        fprintf(stderr, "ehyb_mtmetis: bad header\n");
        return 1;
    }
    uint32_t n = hdr[1], nparts = hdr[2], nthreads = hdr[3];
    uint32_t *xadj = (uint32_t *)malloc(((size_t)n + 1) * sizeof(uint32_t));
    if (!xadj || read_all(f, xadj, ((size_t)n + 1) * sizeof(uint32_t))) return 1;
    size_t nadj = xadj[n];
    uint32_t *adj = (uint32_t *)malloc((nadj ? nadj : 1) * sizeof(uint32_t));
    if (!adj || read_all(f, adj, nadj * sizeof(uint32_t))) return 1;
    ehyb_mtm_wgt *vwgt = NULL, *adjwgt = NULL;
    if (hdr[0] == EHYW_MAGIC) {
        vwgt = (ehyb_mtm_wgt *)malloc((n ? n : 1) * sizeof(ehyb_mtm_wgt));
        adjwgt = (ehyb_mtm_wgt *)malloc((nadj ? nadj : 1) * sizeof(ehyb_mtm_wgt));
        if (!vwgt || !adjwgt || read_all(f, vwgt, (size_t)n * sizeof(ehyb_mtm_wgt)) || read_all(f, adjwgt, nadj * sizeof(ehyb_mtm_wgt))) return 1;
    }
    fclose(f);

    uint32_t *where = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    double *options = mtmetis_init_options();
    options[EHYB_MTMETIS_OPTION_NTHREADS] = (double)nthreads;
    ehyb_mtm_vtx ncon = 1;
    ehyb_mtm_wgt cut = 0;
    int rc = MTMETIS_PartGraphKway(&n, &ncon, xadj, adj, vwgt, NULL, adjwgt, &nparts, NULL, &ub,
                                   options, &cut, where);
    if (rc != EHYB_MTMETIS_SUCCESS) {
        fprintf(stderr, "ehyb_mtmetis: MTMETIS_PartGraphKway returned %d\n", rc);
        return 1;
    }
    f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 1; }
    fwrite(&cut, sizeof cut, 1, f);
    fwrite(where, sizeof(uint32_t), n, f);
    fclose(f);
    return 0;
}
