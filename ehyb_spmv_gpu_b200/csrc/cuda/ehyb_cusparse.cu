/*
 * ehyb_cusparse.cu -- the reference's comparison path, working: CSR SpMV through the cuSPARSE
 * generic API (libehyb_cusparse.so, separate from libehyb.so so that the engine itself has no
 * cuSPARSE dependency).
 *
 * Replaces reference spmv.cu:135-281 (`spmvGeneric`): declared in spmv.h:84-86, never called,
 * built on APIs removed from CUDA 12 (CUSPARSE_CSRMV_ALG1, cusparseCsrmvEx), with fp32 vector
 * descriptors over fp64 data and a (dimension+1)*sizeof(double) row-pointer allocation.  Here:
 * fp64 throughout, CUSPARSE_SPMV_CSR_ALG1 / ALG2 (CUDA 12 enums), cusparseSpMV_preprocess,
 * CUDA-event timing.  This is a measured baseline (bench.py "comparisons"),
 * not part of the product path.
 */
#include <cuda_runtime.h>
#include <cusparse.h>
#include <stdio.h>
#include <stdlib.h>
#include "ehyb_cusparse.h"

#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            snprintf(g_err, sizeof g_err, "%s: %s", #call, cudaGetErrorString(e__));                \
            rc = -3;                                                                                \
            goto done;                                                                              \
        }                                                                                           \
    } while (0)
#define CS(call)                                                                                    \
    do {                                                                                            \
        cusparseStatus_t s__ = (call);                                                              \
        if (s__ != CUSPARSE_STATUS_SUCCESS) {                                                       \
            snprintf(g_err, sizeof g_err, "%s: %s", #call, cusparseGetErrorString(s__));            \
            rc = -3;                                                                                \
            goto done;                                                                              \
        }                                                                                           \
    } while (0)

static char g_err[512];

extern "C" const char *ehyb_cusparse_last_error(void) { return g_err; }

/*
 * y = A x with A = (rowIdx, J, V) of `m` (CSR view of the row-sorted COO, spmv.h:24-29),
 * host vectors; `iters` timed products after `warmup` untimed ones.  alg: 1 or 2
 * (CUSPARSE_SPMV_CSR_ALG1 / ALG2).  *us_per_product (optional) = CUDA-event time per product.
 * Returns 0, or a negative status with ehyb_cusparse_last_error().
 */
extern "C" int ehyb_cusparse_spmv(const matrixCOO *m, const double *x_h, double *y_h, int warmup, int iters, int alg,
                                  float *us_per_product)
{
    int rc = 0;
    g_err[0] = 0;
    if (!m || !x_h || !y_h || iters <= 0 || !m->rowIdx || !m->J || !m->V) {
        snprintf(g_err, sizeof g_err, "ehyb_cusparse_spmv: bad argument");
        return -1;
    }
    const int n = m->dimension;
    const int nnz = m->totalNum;
    int *rowPtr_d = NULL, *col_d = NULL;
    double *val_d = NULL, *x_d = NULL, *y_d = NULL;
    void *buf_d = NULL;
    size_t bufBytes = 0;
    cusparseHandle_t handle = NULL;
    cusparseSpMatDescr_t A = NULL;
    cusparseDnVecDescr_t vx = NULL, vy = NULL;
    cudaEvent_t e0 = NULL, e1 = NULL;
    const double one = 1.0, zero = 0.0;
    const cusparseSpMVAlg_t algo = alg == 2 ? CUSPARSE_SPMV_CSR_ALG2 : CUSPARSE_SPMV_CSR_ALG1;
    float ms = 0.f;

    CK(cudaMalloc(&rowPtr_d, sizeof(int) * ((size_t)n + 1)));
    CK(cudaMalloc(&col_d, sizeof(int) * (size_t)(nnz ? nnz : 1)));
    CK(cudaMalloc(&val_d, sizeof(double) * (size_t)(nnz ? nnz : 1)));
    CK(cudaMalloc(&x_d, sizeof(double) * (size_t)n));
    CK(cudaMalloc(&y_d, sizeof(double) * (size_t)n));
    CK(cudaMemcpy(rowPtr_d, m->rowIdx, sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(col_d, m->J, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(val_d, m->V, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(x_d, x_h, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    CK(cudaMemset(y_d, 0, sizeof(double) * (size_t)n));
    CS(cusparseCreate(&handle));
    CS(cusparseCreateCsr(&A, n, n, nnz, rowPtr_d, col_d, val_d, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F));
    CS(cusparseCreateDnVec(&vx, n, x_d, CUDA_R_64F));
    CS(cusparseCreateDnVec(&vy, n, y_d, CUDA_R_64F));
    CS(cusparseSpMV_bufferSize(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, vx, &zero, vy, CUDA_R_64F, algo, &bufBytes));
    CK(cudaMalloc(&buf_d, bufBytes ? bufBytes : 16));
    CS(cusparseSpMV_preprocess(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, vx, &zero, vy, CUDA_R_64F, algo, buf_d));
    for (int i = 0; i < warmup; ++i)
        CS(cusparseSpMV(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, vx, &zero, vy, CUDA_R_64F, algo, buf_d));
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, 0));
    for (int i = 0; i < iters; ++i)
        CS(cusparseSpMV(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, vx, &zero, vy, CUDA_R_64F, algo, buf_d));
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(y_h, y_d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    if (us_per_product) *us_per_product = ms * 1e3f / (float)iters;
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (vx) cusparseDestroyDnVec(vx);
    if (vy) cusparseDestroyDnVec(vy);
    if (A) cusparseDestroySpMat(A);
    if (handle) cusparseDestroy(handle);
    cudaFree(buf_d); cudaFree(rowPtr_d); cudaFree(col_d); cudaFree(val_d); cudaFree(x_d); cudaFree(y_d);
    return rc;
}

/* reference spmv.h:84-86, with its semantics: MAXIter products of the same x, result in
 * vector_out, the reference's log line.  Errors print and abort (the reference calls exit). */
extern "C" void spmvGeneric(matrixCOO *localMatrix, const double *vector_in, double *vector_out, const int MAXIter)
{
    float us = 0.f;
    if (ehyb_cusparse_spmv(localMatrix, vector_in, vector_out, 10, MAXIter > 0 ? MAXIter : 1, 1, &us) != 0) {
        fprintf(stderr, "spmvGeneric: %s\n", g_err);
        abort();
    }
    printf("iter is %d, cuSPARSE CSR time is %f ms, GPU Gflops is %f\n ", MAXIter, us * 1e-3 * MAXIter,
           2.0 * localMatrix->totalNum / (us * 1e3));
}
