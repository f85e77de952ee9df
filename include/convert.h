/*
 * convert.h -- COO -> EHYB format build in the reference layout (drop-in for reference
 * convert.h:17-20).  The Blackwell-tuned device layout is built by ehyb_layout_build()
 * (ehyb.h); it holds the same logical content and can be mapped back onto this one.
 */
#ifndef CONVERT_H
#define CONVERT_H
#include "spmv.h"
#ifdef __cplusplus
extern "C" {
#endif

/*
 * reference convert.c:316-369.  Allocates (malloc) and fills every host array of
 * outputMatrix exactly as the reference does, byte for byte on
 * every input the reference handles.  Where the reference aborts or is undefined
 *  this implementation defines the behaviour: zero-width slices
 * and matrices without remainder are legal, and rows with more than 512 in-window entries at
 * the head of a partition are stored whole in longVec* (columns from J).
 * Prints the reference's two log lines.  Aborts on inconsistent input.
 */
void COO2EHYB(matrixCOO *inputMatrix, matrixEHYB *outputMatrix, int *sizeBlockELL, int *sizeER);

/* Releases the host arrays allocated by COO2EHYB (the reference leaks them, B-16).
 * partBoundary is aliased from the matrixCOO (reference convert.c:330) and is not freed. */
void EHYBfreeHost(matrixEHYB *m);

#ifdef __cplusplus
}
#endif
#endif
