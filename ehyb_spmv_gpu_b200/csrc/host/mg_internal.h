/* mg_internal.h -- a rank's block of a distributed matrix (shared by mg.c and grid.c). */
#ifndef EHYB_MG_INTERNAL_H
#define EHYB_MG_INTERNAL_H
#include "common.h"

struct ehyb_mg_local {
    int rank, nranks;
    int64_t *rowStarts; /* [nranks+1] */
    int64_t n;          /* local rows */
    int64_t nnz;
    /* local matrix, columns renumbered [own | halo]; natural local row order until finish() */
    int64_t *rowPtr;
    int32_t *col;
    double *val;
    /* halo */
    int64_t nHalo;
    int64_t *haloGlobal; /* [nHalo] sorted global columns */
    int64_t *recvCount;  /* [nranks] */
    /* send */
    int64_t nSend;
    int64_t *sendCount;  /* [nranks] */
    int64_t *sendGlobal; /* [nSend] global rows peers need, grouped by peer */
    int32_t *sendIdx;    /* [nSend] the same as permuted local indices (after finish) */
    /* level 2 */
    matrixCOO coo;       /* permuted local block (after finish; absent for streamed grid blocks) */
    int finished;
    ehyb_layout *layout;
    /* streamed grid blocks (grid.c): no matrixCOO; the level-2 permutation alone */
    int streamed;
    int32_t *perm;       /* [n] local row (level-1 order) -> permuted local row */
    const struct ehyb_grid_decomp *grid; /* borrowed: the decomposition the block was cut from */
};


/* sendIdx[i] = permuted local index of sendGlobal[i] (needs the send list and the permutation) */
int ehyb_mg_local_update_send_idx(ehyb_mg_local *L);

#endif
