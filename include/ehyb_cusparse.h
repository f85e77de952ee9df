/*
 * ehyb_cusparse.h -- the reference's cuSPARSE comparison path (libehyb_cusparse.so).
 *
 * Replaces reference spmv.h:84-86 / spmv.cu:135-281 (`spmvGeneric`: declared, never called,
 * written against cuSPARSE APIs that CUDA 12 removed, fp32 descriptors over fp64 data).  Kept in
 * its own library so that libehyb.so does not depend on cuSPARSE; bench.py reports it as a
 * comparison row measured on the same GPU.  Not part of the product path.
 */
#ifndef EHYB_CUSPARSE_H
#define EHYB_CUSPARSE_H
#include "spmv.h"
#ifdef __cplusplus
extern "C" {
#endif

/* y = A x through cusparseSpMV (CSR, fp64) over the matrix's rowIdx/J/V; host vectors; `iters`
 * timed products after `warmup`; alg 1|2 = CUSPARSE_SPMV_CSR_ALG1|ALG2; *us_per_product
 * (optional) from CUDA events.  0 on success, negative otherwise (ehyb_cusparse_last_error). */
int ehyb_cusparse_spmv(const matrixCOO *m, const double *x_h, double *y_h, int warmup, int iters, int alg,
                       float *us_per_product);
const char *ehyb_cusparse_last_error(void);
/* reference signature and log line (spmv.h:84-86); aborts on error like the reference exits */
void spmvGeneric(matrixCOO *localMatrix, const double *vector_in, double *vector_out, const int MAXIter);

#ifdef __cplusplus
}
#endif
#endif
