/*
 * kernel.h -- per-iteration launchers (drop-in for reference kernel.h:52-60).
 *
 * The reference hard-wires RTX-3090/V100 constants here (kernel.h:20-28: 82/80 SMs,
 * 93 KB shared memory, 1024 threads).  This implementation queries the device at run time
 * (ehyb_device_query) and keeps the old names only as the values used on B200.
 */
#ifndef KERNEL_H
#define KERNEL_H
#include "spmv.h"
#ifdef __cplusplus
extern "C" {
#endif

#define EHYB_WARP 32
#define EHYB_REF_LONG_ROW 512 /* reference threadLongVec, kernel.h:26 */

/*
 * reference kernel.cu:490-518 / :520-552.  inputMatrix must describe DEVICE data prepared by
 * ehyb_upload() (its `b200` field holds the session); vector_in_d / vector_out_d are device
 * pointers of `dimension` doubles, 16-byte aligned.  One y = A*x product, asynchronous on
 * the session's stream.  The `_small` variant is the several-CTAs-per-partition launch
 * (kernelPerPart > 1); both entry points dispatch on the session's plan, and
 * biasIdxBlock_d is accepted for source compatibility and ignored (work is assigned
 * statically).  Abort on a CUDA error.
 */
void matrixVectorEHYB(matrixEHYB *inputMatrix, double *vector_in_d, double *vector_out_d);
void matrixVectorEHYB_small(matrixEHYB *inputMatrix_d, int *biasIdxBlock_d, double *vectorIn_d,
                            double *vectorOut_d);

#ifdef __cplusplus
}
#endif
#endif
