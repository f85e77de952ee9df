/* common.h -- internal helpers of the host side (not installed). */
#ifndef EHYB_COMMON_H
#define EHYB_COMMON_H
#include <stddef.h>
#include <stdint.h>
#include "ehyb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Records a printf-style message for ehyb_last_error() and returns `code`. */
int ehyb_fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
/* Prints the last error and aborts: used by the void-returning drop-in wrappers. */
void ehyb_die(const char *where) __attribute__((noreturn));

/* reorder.c: ehyb_reorder_with_partition for a local block with halo columns [n, ncols) */
int ehyb_reorder_core(matrixCOO *m, const uint32_t *partVec, int ncols);

/* partition.c: one single-threaded mt-metis call in a helper process (re-entrant) */
int ehyb_partition_graph_process(uint32_t n, const uint32_t *xadj, const uint32_t *adjncy, uint32_t nparts, uint32_t *where);

static inline int64_t ehyb_round_up64(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

#ifdef __cplusplus
}
#endif
#endif
