/*
 * spmv.h -- public types and the session entry point of the EHYB SpMV engine.
 *
 * Drop-in for the reference header of the same name (reference spmv.h:7-78): struct and
 * field names, field order and the `spmvGPuEHYB` signature are the reference's.  Unlike the
 * reference header (which only compiles as C++: `bool` without <stdbool.h>, a default member
 * initialiser at spmv.h:58, `extern "C"` without a guard) this one is real C11 and is
 * also usable from C++.  New fields are appended after the reference's, never inserted.
 */
#ifndef SPMV_H
#define SPMV_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference spmv.h:7-15 */
typedef struct _cb {
    bool PRECOND;
    bool GPU;
    bool RODR;
    bool CACHE;
    bool BLOCK;
    bool FACT;
    bool SORT;
} cb_s;

/* reference spmv.h:17-33.  Row-sorted COO + CSR pointer + partition parameters. */
typedef struct _matrixCOO {
    int totalNum;             /* nnz */
    int dimension;            /* n */
    int maxCol;               /* longest row */
    int nParts;               /* number of partitions P */
    uint16_t vectorCacheSize; /* x window length W (elements) */
    int16_t kernelPerPart;    /* CTAs per partition */
    int *rowIdx;              /* [n+1] */
    int *numInRow;            /* [n] row lengths */
    int *numInRow2;           /* [n] entries of the row whose column is in its partition's window */
    int *I;                   /* [nnz] */
    int *J;                   /* [nnz] */
    double *V;                /* [nnz] */
    double *diag;             /* [n] */
    int *partBoundary;        /* [nParts+1] first row of every partition (permuted numbering) */
    int *reorderList;         /* [n] old row -> new row */
} matrixCOO;

/* reference spmv.h:35-63: the EHYB arrays in the reference layout.
 * The same struct describes host arrays (after COO2EHYB) and device arrays (after
 * ehyb_upload_reference_layout / inside spmvGPuEHYB). */
typedef struct _matrixEHYB {
    int dimension;
    int nParts;
    int16_t vectorCacheSize;
    int kernelPerPart;
    int numOfRowER;
    int *warpIdxER_d;
    int *reorderList;
    int *reorderListER;
    int16_t *widthVecBlockELL;
    int *biasVecBlockELL;
    int16_t *colBlockELL;
    double *valBlockELL;
    int *partBoundary;
    int16_t *widthVecER;
    int *rowVecER;
    int *biasVecER;
    int *colER;
    double *valER;
    double *outER;
    /* long rows (CSR-like, whole rows, columns are global permuted indices) */
    int nLongVec;
    int *longVecBoundary;
    int *longVecRow;
    int *longVecCol;
    double *longVecVal;
    /* ---- appended by this implementation ---- */
    void *b200; /* device session (struct ehyb_handle*) when this struct describes device data */
} matrixEHYB;

/* reference spmv.h:65-73 */
static inline void init_cb(cb_s *in_s)
{
    in_s->PRECOND = false;
    in_s->GPU = false;
    in_s->RODR = true;
    in_s->BLOCK = true;
    in_s->CACHE = true;
    in_s->FACT = true;
    in_s->SORT = false;
}

/*
 * reference spmv.h:75-78 / spmv.cu:61-133.  One SpMV session: localMatrix is the matrix after
 * matrixReorder[_unsym]; vectorIn / vectorOut are the permuted x and y (host).  Builds the
 * device format, uploads, runs 10 warm-up products and MAXIter timed products of the same
 * x, downloads y and prints the reference's report line plus the roofline lines.
 * *realIter receives the number of timed products (the reference never writes it).
 * Aborts the process on a CUDA or format error (the reference ignores them).
 */
void spmvGPuEHYB(matrixCOO *localMatrix, const double *vectorIn, double *vectorOut,
                 const int MAXIter, int *realIter);
/* The same session on an existing tuned layout (struct ehyb_layout of ehyb.h: built by the
 * caller or loaded from the binary cache): everything of spmvGPuEHYB after COO2EHYB. */
struct ehyb_layout;
void spmvGPuEHYB_layout(const struct ehyb_layout *layout, const double *vectorIn, double *vectorOut,
                        const int MAXIter, int *realIter);


#ifdef __cplusplus
}
#endif
#endif
