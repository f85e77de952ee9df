/*
 * reordering.h -- host partition + reorder stage (drop-in for reference reordering.h:6-10).
 */
#ifndef REORDERING_H
#define REORDERING_H
#include "spmv.h"
#ifdef __cplusplus
extern "C" {
#endif

/*
 * reference reordering.c:231-378 (symmetric: graph = the matrix pattern incl. self loops,
 * mt-metis with 1 thread) and :41-228 (unsymmetric: pattern of A+A^T, 6 threads).
 * In : the reader's matrixCOO (I/J/V, rowIdx, numInRow, nParts, vectorCacheSize; numInRow2
 *      zeroed).  Out: permuted row-sorted I/J/V (the caller's arrays are freed and replaced,
 *      as in the reference), rowIdx, numInRow, numInRow2 (in-window counts), partBoundary
 *      (re-allocated), reorderList.  Aborts on failure.
 */
void matrixReorder(matrixCOO *localMatrixCOO);
void matrixReorder_unsym(matrixCOO *localMatrixCOO);

/* reference reordering.c:380-384: v_rodr[rodr_list[i]] = v_in[i] */
void vectorReorder(const int dimension, const double *v_in, double *v_rodr, const int *rodr_list);
/* reference reordering.c:386-391: v[i] = v_rodr[rodr_list[i]] */
void vectorRecover(const int dimension, const double *v_rodr, double *v, const int *rodr_list);

#ifdef __cplusplus
}
#endif
#endif
