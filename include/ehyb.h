/*
 * ehyb.h -- C ABI of the B200-native EHYB SpMV engine (libehyb.so).
 *
 * Everything here is new relative to the reference; the reference's own entry points stay
 * available under their old names (spmv.h, kernel.h, convert.h, reordering.h, mmio.h) and are
 * thin wrappers over this layer.  Conventions: plain C, plain pointers and sizes; every
 * function returns EHYB_OK (0) or a negative ehyb_status and never calls exit();
 * ehyb_last_error() gives the message of the last failure on the calling thread.
 *
 * Which reference interface each group replaces:
 *   device / plan      solver_test.c:158-182, :53-77 (partition-parameter heuristic built on
 *                      kernel.h:20-25 compile-time constants)
 *   graph / reorder    reordering.c:41-228, :231-378 (split so that the partition vector can
 *                      be injected; matrixReorder* are wrappers)
 *   layout             convert.c:316-369 (COO2EHYB) - emits the Blackwell-tuned layout
 *   session            spmv.cu:6-60 (cudaMallocTransDataEHYB), spmv.cu:61-133 (spmvGPuEHYB),
 *                      kernel.cu:324-380, :490-518 (matrixVectorEHYB launchers)
 *   multi-GPU          none (new layer)
 */
#ifndef EHYB_H
#define EHYB_H

#include <stddef.h>
#include <stdint.h>
#include "spmv.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ehyb_status {
    EHYB_OK = 0,
    EHYB_ERR_ARG = -1,       /* invalid argument / inconsistent input */
    EHYB_ERR_NOMEM = -2,     /* host allocation failed */
    EHYB_ERR_CUDA = -3,      /* a CUDA call failed, or no CUDA device */
    EHYB_ERR_PARTITION = -4, /* the partitioner failed or is not available */
    EHYB_ERR_IO = -5,        /* file problem */
    EHYB_ERR_LIMIT = -6,     /* a format limit was exceeded (e.g. window larger than 65536) */
    EHYB_ERR_NCCL = -7,      /* an NCCL call failed */
    EHYB_ERR_PEER = -8       /* peer-memory exchange: no P2P access, or a neighbour never delivered */
} ehyb_status;

const char *ehyb_last_error(void);
const char *ehyb_version(void);
/* threads of the host-side OpenMP code (format build, reorder, generators); n <= 0: unchanged.
 * Process launchers tend to export OMP_NUM_THREADS=1 to their workers. */
void ehyb_set_host_threads(int n);
int ehyb_get_host_threads(void);

/* ------------------------------------------------------------------------------------ */
/* device query and partition-parameter plan                                              */
/* ------------------------------------------------------------------------------------ */

typedef struct ehyb_device_info {
    int device;
    int sm_count;             /* 148 on B200 */
    int smem_optin_bytes;     /* max dynamic shared memory per CTA (232448 on B200) */
    int smem_per_sm_bytes;    /* 233472 on B200 */
    int l2_bytes;
    int cc_major, cc_minor;
    int max_persist_l2_bytes; /* cudaDevAttrMaxPersistingL2CacheSize */
    size_t hbm_bytes;
    char name[64];
} ehyb_device_info;

int ehyb_device_count(int *count);
int ehyb_device_query(int device, ehyb_device_info *out);
/* Nominal B200 values, so that host-side planning and format build can run without a GPU. */
void ehyb_device_info_b200(ehyb_device_info *out);

#define EHYB_KERNEL_DIRECT 1
#define EHYB_KERNEL_STAGED 2
#define EHYB_KERNEL_PERSISTENT 3

typedef struct ehyb_plan_t {
    int nParts;      /* P */
    int W;           /* x window length in elements (vectorCacheSize) */
    int ctasPerPart; /* kernelPerPart */
    int threads;     /* threads per CTA */
    int ctasPerSM;   /* resident CTAs per SM the plan was sized for */
} ehyb_plan_t;

/* B200 plan from the matrix size and the device (replaces solver_test.c:158-182). */
int ehyb_plan(int n, const ehyb_device_info *dev, ehyb_plan_t *out);
/* The same for a given kernel (EHYB_KERNEL_* below; ehyb_plan = the staged kernel, which the
 * multi-GPU sessions use): EHYB_KERNEL_PERSISTENT sizes the partitions for the persistent kernel
 * of single-GPU sessions (three or more smaller partitions per SM), see plan.c. */
int ehyb_plan_kernel(int n, const ehyb_device_info *dev, int kernel, ehyb_plan_t *out);
/* What the reference's reader does in solver_test.c:158-182 / :53-77 (pick W, nParts, kernelPerPart),
 * for a B200: the plan and the kernel (*kernel, may be NULL) a driver should use for n rows and nnz entries: one
 * partition per SM (staged kernel) when matrix + vectors fit 3/4 of L2, the persistent kernel's plan up
 * to ~40 entries per row, the staged plan beyond (plan.c). */
int ehyb_plan_auto(int n, int64_t nnz, const ehyb_device_info *dev, ehyb_plan_t *out, int *kernel);
/* The reference's own heuristic for an 82/80-SM, 93 KB device, including its int16_t wrap
 * (SURVEY.md Appendix C).  ctasPerPart = 0 where the reference leaves it uninitialised. */
int ehyb_plan_reference(int n, int symmetric, ehyb_plan_t *out);

/* ------------------------------------------------------------------------------------ */
/* graph, partition, reorder                                                              */
/* ------------------------------------------------------------------------------------ */

/* Graph handed to the partitioner.  symmetric: xadj = rowIdx, adjncy = J (self loops kept),
 * reordering.c:239-264.  Otherwise the pattern of A + A^T with duplicates kept,
 * reordering.c:50-89.  *xadj / *adjncy are malloc'd; free with ehyb_free_host. */
int ehyb_build_graph(const matrixCOO *m, int symmetric, uint32_t **xadj, uint32_t **adjncy);

/* Partitioner hook.  The default runs the pinned mt-metis binary (third_party/mtmetis) out of
 * process through bin/ehyb_mtmetis (path: $EHYB_MTMETIS_BIN, else next to libehyb.so);
 * bin/spmv.out installs an in-process call.  ubvec 1.001, ncon 1, no weights, as the
 * reference (reordering.c:270-293). */
typedef int (*ehyb_partition_fn)(uint32_t nvtxs, const uint32_t *xadj, const uint32_t *adjncy,
                                 uint32_t nparts, uint32_t nthreads, float ubvec, uint32_t *where,
                                 void *user);
void ehyb_set_partitioner(ehyb_partition_fn fn, void *user);
int ehyb_partition_graph(uint32_t nvtxs, const uint32_t *xadj, const uint32_t *adjncy,
                         uint32_t nparts, uint32_t nthreads, uint32_t *where);

/* Deterministic and parallel partition stage (csrc/host/hierpart.c): the rows are
 * contracted in blocks, ONE single-threaded mt-metis call cuts the block graph into `pieces`
 * pieces, and every piece is partitioned on its own by a single-threaded mt-metis process, all at
 * once.  Same partition vector on every run (the reference's threaded call, reordering.c:120, is
 * not), wall time of the largest piece.  pieces <= 1 or n < 65536: the reference's single call.
 * ehyb_set_partition_pieces(T) makes ehyb_reorder / matrixReorder* use it ($EHYB_PARTITION_PIECES
 * overrides; 0 = the reference's call, the default). */
int ehyb_partition_graph_hier(uint32_t nvtxs, const uint32_t *xadj, const uint32_t *adjncy, uint32_t nparts, int pieces,
                              uint32_t *where);
void ehyb_set_partition_pieces(int pieces);
int ehyb_get_partition_pieces(void);

/* Everything of matrixReorder after the mt-metis call (reordering.c:299-377), for a given
 * partition vector partVec[old row] in [0, nParts).  Same in/out contract as matrixReorder. */
int ehyb_reorder_with_partition(matrixCOO *m, const uint32_t *partVec);
/* Status-returning matrixReorder / matrixReorder_unsym. */
int ehyb_reorder(matrixCOO *m, int symmetric);
/* Contiguous-block partition vector (rows [i*n/P, (i+1)*n/P) -> part i): the structured-grid
 * alternative used for per-GPU slabs and very large grids where mt-metis is impractical. */
int ehyb_partition_blocks(uint32_t n, uint32_t nparts, uint32_t *where);

void ehyb_free_host(void *p);

/* ------------------------------------------------------------------------------------ */
/* Blackwell-tuned device layout (host side)                                              */
/* ------------------------------------------------------------------------------------ */

#define EHYB_SLICE_ROWS 64 /* rows per slice: lane l owns rows l and l+32 of the slice */

typedef struct ehyb_layout ehyb_layout;

typedef struct ehyb_layout_opts {
    int W;                  /* x window length; 0 = matrixCOO.vectorCacheSize */
    int ctasPerPart;        /* 0 = matrixCOO.kernelPerPart (min 1) */
    double er_fill;         /* in-slice remainder column kept while >= er_fill*64 rows use it;
                               0 = keep every remainder entry in its slice; < 0 = choose between 0
                               and 0.5 by stored bytes (the default of spmvGPuEHYB) */
    int long_row_threshold; /* rows at a partition head with more in-window entries than this go
                               whole to the overflow list; 0 = min(512 (reference threadLongVec),
                               max(32, twice a warp's fair share of the partition's columns)): a
                               slice is walked by one warp, see layout.c */
    int64_t ncols;          /* columns of the local operator (n + halo); 0 = n */
    int halo_in_overflow;   /* entries with a halo column (>= n) always go to the overflow list,
                               which the multi-GPU product runs after the halo exchange */
    int cache_cap;          /* remainder cache: at most this many columns outside the window are
                               cached per partition (0 = what fits a B200 CTA next to the window
                               and ~100 KB of staging, at least 4096; < 0 = none: all remainder
                               entries go to the overflow list) */
    double min_coverage;    /* if fewer than this fraction of the entries would live in slices
                               (window + remainder cache), every entry goes to the overflow list
                               and the product is memset + overflow kernel; 0 = 0.2, < 0 = never */
} ehyb_layout_opts;

#define EHYB_DEFAULT_CACHE_CAP 4096
#define EHYB_DEFAULT_MIN_COVERAGE 0.2

typedef struct ehyb_slice_desc {
    uint32_t off256; /* byte offset of the slice in the blob / 256 */
    uint16_t w;      /* ELL width (columns), 16-bit window-local indices */
    uint16_t wr;     /* in-slice remainder width, 16-bit indices into the partition's cache */
} ehyb_slice_desc;

typedef struct ehyb_part_desc {
    int32_t rowStart, rowEnd;       /* permuted rows of the partition */
    int32_t sliceStart, sliceEnd;   /* its slices */
    int32_t cacheStart, cacheCount; /* its remainder cache list in cacheCols */
    int32_t reserved[2];
} ehyb_part_desc;

/* Read-only view of a built layout (pointers stay owned by the layout). */
typedef struct ehyb_layout_view {
    int64_t n, ncols, nnz;
    int32_t nParts, W, ctasPerPart, nSlices;
    const ehyb_part_desc *parts;   /* [nParts] */
    const ehyb_slice_desc *slices; /* [nSlices] */
    const unsigned char *blob;     /* [blobBytes] slice data, see DESIGN.md "data layout" */
    int64_t blobBytes;
    int64_t nOverflow;             /* overflow (COO, row-sorted, entry order kept) */
    const int32_t *ovfRow, *ovfCol;
    const double *ovfVal;
    const int32_t *cacheCols;      /* [cacheTotal] per-partition remainder cache lists */
    int64_t cacheTotal;
    int32_t cacheMax;              /* longest list: sizes the shared-memory cache */
    int32_t haloInOverflow;        /* built with halo_in_overflow (no halo column in a cache list) */
    /* statistics */
    int64_t nnzEll, nnzRemInSlice, nnzOverflow, padEll, padRem, nLongRows;
    int64_t algBytes;              /* 8 nnz + 2 nnzEll + 4 (nnz - nnzEll) + 16 n, BASELINE.md sec. 3 */
    int64_t formatBytes;           /* blob + descriptors + overflow arrays actually stored */
} ehyb_layout_view;

/* Build from a reordered matrixCOO (after matrixReorder*, or any row-sorted COO with
 * partBoundary/nParts set).  Classification of entries is the reference's
 * (convert.c:247-267): entry -> ELL iff partStart <= J < partStart+W. */
int ehyb_layout_build(const matrixCOO *m, const ehyb_layout_opts *opts, ehyb_layout **out);
/* Same from raw CSR arrays with 64-bit row pointers (local blocks of the multi-GPU path). */
int ehyb_layout_build_csr(int64_t n, const int64_t *rowPtr, const int32_t *col, const double *val,
                          int nParts, const int32_t *partBoundary, const ehyb_layout_opts *opts,
                          ehyb_layout **out);
int ehyb_layout_get(const ehyb_layout *L, ehyb_layout_view *view);
/* De-interleave the tuned layout back to the reference layout: allocates and
 * fills `out` like COO2EHYB would for the same matrix (requires W % 32 == 0, W < 32768 and
 * sizes below 2^31).  Used by the parity tests; the device never sees this form. */
int ehyb_layout_to_reference(const ehyb_layout *L, matrixEHYB *out, int *sizeBlockELL, int *sizeER);
void ehyb_layout_free(ehyb_layout *L);

/* Streamed build (BASELINE.json config 5: 3.6 G entries never exist as one COO): the permuted
 * matrix arrives a few consecutive partitions at a time - rows [partBoundary[0],
 * partBoundary[nParts]) as CSR with rowPtr relative to the chunk (rowPtr[0] == 0) and absolute
 * local column numbers - and the slices are appended to one blob; peak memory = the layout + one
 * chunk.  Same layout as ehyb_layout_build_csr of the whole matrix, byte for byte, for
 * er_fill >= 0 (er_fill < 0 decides per chunk); opts->W must be set; the all-overflow fallback
 * (min_coverage) does not apply.  Layouts of more than 64 M rows drop the per-row bookkeeping
 * that only ehyb_layout_to_reference and the cache file need. */
typedef struct ehyb_layout_builder ehyb_layout_builder;
int ehyb_layout_builder_begin(int64_t n, const ehyb_layout_opts *opts, ehyb_layout_builder **out);
int ehyb_layout_builder_add(ehyb_layout_builder *B, int nParts, const int32_t *partBoundary, const int64_t *rowPtr,
                            const int32_t *col, const double *val);
int ehyb_layout_builder_finish(ehyb_layout_builder *B, ehyb_layout **out); /* releases B */
void ehyb_layout_builder_abort(ehyb_layout_builder *B);

/* Binary cache (the reference re-runs reader, mt-metis, reorder and COO2EHYB on
 * every invocation).  ehyb_layout_save/load move a layout alone.  ehyb_cache_save/load move the
 * whole result of the pipeline for a source file: the layout plus - all optional together - the
 * permutation reorderList[n], x[n], the golden product y_golden[n] and absAx[n] = |A||x| (for
 * the accuracy gate).  The cache records size + mtime of source_path and the partition
 * parameters; ehyb_cache_load fails with EHYB_ERR_IO when the file is absent, truncated, fails
 * its checksum, was built from another version of the source, or - with plan != NULL - for
 * other parameters.  Output arrays are malloc'd (ehyb_free_host). */
int ehyb_layout_save(const ehyb_layout *L, const char *path);
int ehyb_layout_load(const char *path, ehyb_layout **out);
/* Options that shaped the layout but are not partition parameters (er_fill, cache_cap, the partition
 * stage, ...), folded by the caller into one tag: a cache written under another tag is rejected.
 * Process-wide; 0 = none (the default). */
void ehyb_cache_set_options_tag(uint64_t tag);
int ehyb_cache_save(const char *path, const char *source_path, const ehyb_layout *L, int symmetric, const int *reorderList,
                    const double *x, const double *y_golden, const double *absAx);
int ehyb_cache_load(const char *path, const char *source_path, const ehyb_plan_t *plan, ehyb_layout **L, int *n,
                    int *symmetric, int **reorderList, double **x, double **y_golden, double **absAx);

/* ------------------------------------------------------------------------------------ */
/* device session                                                                         */
/* ------------------------------------------------------------------------------------ */

typedef struct ehyb_handle ehyb_handle;

typedef struct ehyb_session_opts {
    int device;         /* CUDA device ordinal */
    int threads;        /* threads per CTA, 0 = default */
    int use_graph;      /* capture the per-product launches in a CUDA graph (default 1) */
    int l2_persist_x;   /* L2 access-policy window on x for the remainder gathers (default 1) */
    int64_t halo_cols;  /* extra x entries after the n local ones (multi-GPU), default 0 */
    int kernel;         /* 0 = default (persistent where the layout allows it, else staged),
                           1 = direct (matrix streamed with 128-bit global loads),
                           2 = staged (matrix streamed through shared memory by TMA),
                           3 = persistent (staged, one CTA per SM over several partitions with
                           the window + remainder cache double-buffered; single GPU) */
} ehyb_session_opts;


void ehyb_session_opts_default(ehyb_session_opts *o);

/* Upload a layout: one allocation + one H2D per array, persistent buffers, own stream. */
int ehyb_upload(const ehyb_layout *L, const ehyb_session_opts *opts, ehyb_handle **out);
/* y = A x with device vectors (16-byte aligned, x has ncols entries); asynchronous on the
 * session stream. */
int ehyb_spmv(ehyb_handle *h, const double *x_d, double *y_d);
/* Same with HOST vectors: H2D of x, product, D2H of y, synchronous.  This is the call the
 * reference's spmvGPuEHYB makes around its timed loop (spmv.cu:109-117). */
int ehyb_spmv_host(ehyb_handle *h, const double *x_h, double *y_h);
/* Pipelined stream of host products: for i in [0,count): y_h[i] = A x_h[i], with the copies
 * of product i+1 overlapping the kernels of product i (pinned staging, 2 copy streams). */
int ehyb_spmv_host_batch(ehyb_handle *h, const double *const *x_h, double *const *y_h, int count);
/* Device buffers owned by the session (x: ncols doubles, y: n doubles). */
int ehyb_session_vectors(ehyb_handle *h, double **x_d, double **y_d);
int ehyb_set_x(ehyb_handle *h, const double *x_h);
int ehyb_get_y(ehyb_handle *h, double *y_h);
/* `iters` products of the session's own x into its own y between two CUDA events on the
 * session stream, after `warmup` untimed ones.  *ms_total = elapsed milliseconds.
 * If kernel_ms != NULL it receives the summed duration of the main kernel alone, measured
 * with per-launch events in a second pass of the same length. */
/* y = A x and *dot_d += x . y in ONE launch (p.Ap of a CG iteration - the caller the reference's unused
 * CG helpers imply, kernel.cu:13-42, :288-321; dot_d in device memory, not zeroed here).  ehyb_spmv_dot_supported: 1 where the product is a single staged / persistent launch. */
int ehyb_spmv_dot_supported(const ehyb_handle *h);
int ehyb_spmv_dot(ehyb_handle *h, const double *x_d, double *y_d, double *dot_d);
int ehyb_spmv_dot_host(ehyb_handle *h, const double *x_h, double *y_h, double *dot_h); /* host vectors, synchronous */
int ehyb_time_spmv(ehyb_handle *h, int warmup, int iters, float *ms_total, float *kernel_ms);
/* The same with a COLD L2 (L2-resident matrices are reported with and without a flush; the
 * reference times spmv.cu:72-101 warm only): `flush_bytes` of scratch memory (>= 2 x the L2 size) are overwritten before
 * every product, each product is timed by its own event pair; *ms_sum = their sum over `iters`. */
int ehyb_time_spmv_flushed(ehyb_handle *h, int warmup, int iters, size_t flush_bytes, float *ms_sum);
/* Number of kernel launches one product issues (1, or 2 when the overflow list is not empty). */
int ehyb_launches_per_spmv(const ehyb_handle *h);
int ehyb_sync(ehyb_handle *h);
void *ehyb_stream(ehyb_handle *h); /* cudaStream_t */
void ehyb_free(ehyb_handle *h);

/* Development aid: with EHYB_TRACE=1 in the environment when the session is created, the main
 * kernel records a per-CTA timeline of the last product (8 words per CTA, see
 * ehyb_device.cu); *ctas = 0 when tracing is off.  out may be NULL to query the size. */
int ehyb_trace_read(ehyb_handle *h, unsigned long long *out, int *ctas);

/* Device-side description in the reference's struct (for matrixVectorEHYB callers): fills
 * `d` with dimension/nParts/... and d->b200 = h.  No arrays in the reference layout exist
 * on the device. */
int ehyb_describe(ehyb_handle *h, matrixEHYB *d);

/* ------------------------------------------------------------------------------------ */
/* solver shell on top of the product: Jacobi-preconditioned CG,   */
/* the loop the reference's unused helpers outline (kernel.cu:13-42, :288-321; cb_s.PRECOND; */
/* matrixCOO.diag; the never-written realIter of spmvGPuEHYB)                                */
/* ------------------------------------------------------------------------------------ */

typedef struct ehyb_pcg_opts {
    int max_iters;   /* default 1000 */
    double rtol;     /* stop when |r| <= rtol |b| (recurrence residual), default 1e-10 */
    int check_every; /* iterations between two looks at the residual from the host, default 8:
                        the scalars of the iteration live on the device */
} ehyb_pcg_opts;

typedef struct ehyb_pcg_result {
    int iters;                /* products performed (a multiple of check_every unless max_iters stopped it) */
    int converged;
    double rel_residual;      /* |r| / |b| of the recurrence at the last check */
    double true_rel_residual; /* |b - A x| / |b|, one extra product at the end */
    float ms;                 /* CUDA-event time of the iterations on the session stream */
} ehyb_pcg_result;

void ehyb_pcg_opts_default(ehyb_pcg_opts *o);
/* Solves A x = b for the symmetric positive definite matrix of session h.  Host vectors in the
 * session's (permuted) numbering, n entries; x starts from zero.  diag_h = the matrix diagonal
 * in the same numbering for Jacobi preconditioning, or NULL for plain CG. */
int ehyb_pcg_solve(ehyb_handle *h, const double *diag_h, const double *b_h, double *x_h, const ehyb_pcg_opts *opts,
                   ehyb_pcg_result *res);
/* CUDA device of the session */
int ehyb_session_device(const ehyb_handle *h);
/* name of the kernel that computes the session's products */
const char *ehyb_session_kernel(const ehyb_handle *h);
/* rows and columns of the session's operator */
int ehyb_session_size(const ehyb_handle *h, int64_t *n, int64_t *ncols);

/* ------------------------------------------------------------------------------------ */
/* multi-GPU: one process per GPU, rows distributed in contiguous blocks, x halo exchanged    */
/* every product (no reference counterpart).  Two exchanges:             */
/*   EHYB_MG_P2P   (product) the main kernel itself stores the x entries its neighbours need   */
/*                 into their memory over NVLink (CUDA IPC mappings + epoch flags) and serves  */
/*                 halo columns from the shared-memory remainder cache: one launch per product */
/*   EHYB_MG_NCCL  (baseline) pack kernel + grouped ncclSend/ncclRecv on a second stream,      */
/*                 overlapped with the main kernel; halo entries live in the overflow list     */
/* ------------------------------------------------------------------------------------ */

typedef struct ehyb_mg_local ehyb_mg_local;     /* host: a rank's block, halo and send lists */
typedef struct ehyb_mg_session ehyb_mg_session; /* device: the block on its GPU + its exchange state */

#define EHYB_MG_NCCL 0
#define EHYB_MG_P2P 1

/* rowStarts[nranks+1]: global row range of every rank.  rowPtr/colGlobal/val: the rank's rows
 * (CSR, global column indices).  Computes the halo (sorted external columns, grouped by owner). */
int ehyb_mg_local_build(int rank, int nranks, const int64_t *rowStarts, const int64_t *rowPtr,
                        const int64_t *colGlobal, const double *val, ehyb_mg_local **out);
/* What this rank receives: haloGlobal[nHalo] (sorted; halo column k is local column n+k) and
 * recvCount[nranks].  Peer g must be told the slice of haloGlobal it owns (any transport). */
int ehyb_mg_local_halo(const ehyb_mg_local *L, int64_t *nHalo, const int64_t **haloGlobal,
                       const int64_t **recvCount);
/* What peers asked from this rank: sendCount[nranks] and the concatenated global rows. */
int ehyb_mg_local_set_send(ehyb_mg_local *L, const int64_t *sendCount, const int64_t *sendGlobal);
/* Graph of the block's own columns for the level-2 partitioner (malloc'd, ehyb_free_host). */
int ehyb_mg_local_graph(const ehyb_mg_local *L, uint32_t **xadj, uint32_t **adjncy);
/* Level-2 partition (partVec[n_local], NULL = contiguous blocks), permutation, tuned layout.
 * exchange = EHYB_MG_NCCL: every halo entry in the overflow list (the main kernel must not
 * depend on the exchange); EHYB_MG_P2P: halo columns are ordinary remainder columns. */
int ehyb_mg_local_finish(ehyb_mg_local *L, int nParts, int W, int ctasPerPart, const uint32_t *partVec,
                         double er_fill, int exchange);
int ehyb_mg_local_view(const ehyb_mg_local *L, const matrixCOO **coo, const ehyb_layout **layout,
                       int64_t *nSend, const int32_t **sendIdx, const int64_t **sendCount);
void ehyb_mg_local_free(ehyb_mg_local *L);

/* 128-byte NCCL unique id (rank 0 creates it, the caller distributes it). */
int ehyb_mg_unique_id(void *id128);
/* NCCL exchange.  Collective: uploads the block and joins the communicator. */
int ehyb_mg_session_create(const ehyb_mg_local *L, int rank, int nranks, int device, const void *id128,
                           ehyb_mg_session **out);
/* Peer-memory exchange, three steps (the caller moves the blobs between the ranks, any
 * transport):  create_p2p (uploads the block, allocates halo buffers + flags)  ->  p2p_export
 * (EHYB_MG_P2P_BLOB_BYTES describing this rank's buffers)  ->  p2p_connect with the blobs of
 * ALL ranks (rank-major) and recvOffsetOnPeer[g] = position of this rank's first entry in rank
 * g's halo list (= sum of g's recvCount[0..rank)).  Needs P2P access between the GPUs
 * (NVLink/NVSwitch) and nranks <= 32. */
#define EHYB_MG_P2P_BLOB_BYTES 128
int ehyb_mg_p2p_supported(int device, int nranks, int *supported);
int ehyb_mg_session_create_p2p(const ehyb_mg_local *L, int rank, int nranks, int device, ehyb_mg_session **out);
int ehyb_mg_p2p_export(ehyb_mg_session *s, void *blob);
int ehyb_mg_p2p_connect(ehyb_mg_session *s, const void *blobs, const int64_t *recvOffsetOnPeer);
/* The same connection when all the ranks live in ONE process (the C driver, bin/spmv.out -G N: one
 * host thread per GPU): plain peer access instead of CUDA IPC; sessions[r] = rank r's session. */
int ehyb_mg_p2p_connect_local(ehyb_mg_session *const *sessions, int nranks);
/* Set when a wait on a neighbour ran into the time limit ($EHYB_P2P_TIMEOUT_MS, default 10 s)
 * since the session was created: the products since then are not valid. */
int ehyb_mg_status(ehyb_mg_session *s, int *timed_out);
int ehyb_mg_session_handle(ehyb_mg_session *s, ehyb_handle **h);
/* y_local = A_block [x_local | halo].  x_d: the n local entries (NCCL exchange: n + nHalo, the
 * halo part is filled here).  Collective in the sense that every rank has to call it the same
 * number of times; asynchronous on the session stream. */
int ehyb_mg_spmv(ehyb_mg_session *s, double *x_d, double *y_d);
int ehyb_mg_time_spmv(ehyb_mg_session *s, int warmup, int iters, float *ms_total);
/* Pipelined stream of distributed products with HOST vectors (the n local entries of x and y, in
 * pinned memory for true overlap): y_h[i] = A_block [x_h[i] | halo]; the copies of products i+1
 * and i-1 overlap the kernel of product i, the halo travels between the GPUs inside the products.
 * Collective: every rank calls it with the same count. */
int ehyb_mg_spmv_host_batch(ehyb_mg_session *s, const double *const *x_h, double *const *y_h, int count);
/* Kernel launches one distributed product issues. */
int ehyb_mg_launches_per_spmv(const ehyb_mg_session *s);
/* In-stream all-reduce (sum over the ranks) of count <= 4 doubles at vals_d, on the session stream:
 * what a distributed solver needs for its dot products.  Collective.  Peer-memory sessions exchange
 * the partial sums through a mailbox every rank maps and add them in rank order (the same bits on
 * every GPU); NCCL sessions call ncclAllReduce. */
int ehyb_mg_allreduce_sum(ehyb_mg_session *s, double *vals_d, int count);
/* ehyb_mg_spmv with the fused dot of ehyb_spmv_dot: *dot_d += x_own . y over THIS rank's rows. */
int ehyb_mg_spmv_dot_supported(const ehyb_mg_session *s);
int ehyb_mg_spmv_dot(ehyb_mg_session *s, double *x_d, double *y_d, double *dot_d);
int ehyb_mg_session_ranks(const ehyb_mg_session *s, int *rank, int *nranks);
/* Distributed (preconditioned) conjugate gradients: ehyb_pcg_solve over the ranks' blocks (no reference
 * counterpart; its cb_s.PRECOND / realIter hooks, spmv.h:7-15, :75-78, are what it fills in).  Collective;
 * every rank passes ITS rows of b, of the diagonal and of x (the block's permuted numbering, n local
 * entries).  The dot products are summed over the GPUs by ehyb_mg_allreduce_sum, in rank order: all
 * ranks see the same residuals and stop in the same iteration. */
int ehyb_mg_pcg_solve(ehyb_mg_session *s, const double *diag_h, const double *b_h, double *x_h, const ehyb_pcg_opts *opts,
                      ehyb_pcg_result *res);
/* Every rank must have finished its products before any rank frees its session (barrier). */
void ehyb_mg_session_free(ehyb_mg_session *s);
/* Rows of z-planes [z0, z1) of the 27-point stencil on nx x ny x nz (full rows, global
 * columns ascending): a rank's slab of BASELINE.json config 5, generated in place. */
int ehyb_gen_stencil27_rows(int nx, int ny, int64_t nz, int64_t z0, int64_t z1, int64_t **rowPtr,
                            int64_t **col, double **val);

/* ---- BASELINE.json config 5: the 27-point stencil on nx x ny x nz, sharded (csrc/host/grid.c) ----
 * The grid is cut into bricks of bx x by x bz cells.  Level 1 assigns bricks to GPUs (owner[]:
 * the mt-metis k = nranks partition of ehyb_grid_brick_graph - the call of reordering.c:270-293
 * on the coarsened graph, with vertex and edge weights - or NULL for contiguous runs of bricks);
 * level 2: every brick is one EHYB partition.  ehyb_mg_grid_build streams a rank's block into the
 * tuned layout without ever holding its COO (peak: the layout + 4 B per row + one chunk) and
 * returns it finished: ehyb_mg_local_halo -> exchange -> ehyb_mg_local_set_send -> session.
 * Level-1 ids (what the halo and send lists carry) are rank-major, then brick, then cell. */
typedef struct ehyb_grid_decomp ehyb_grid_decomp;
int ehyb_grid_brick_graph(int nx, int ny, int nz, int bx, int by, int bz, int64_t *nBricks, uint32_t **xadj, uint32_t **adjncy,
                          int32_t **vwgt, int32_t **adjwgt);
/* weighted k-way partition by the pinned mt-metis binary (through bin/ehyb_mtmetis) */
int ehyb_partition_graph_weighted(uint32_t nvtxs, const uint32_t *xadj, const uint32_t *adjncy, const int32_t *vwgt,
                                  const int32_t *adjwgt, uint32_t nparts, uint32_t nthreads, float ubvec, uint32_t *where);
int ehyb_grid_decomp_create(int nx, int ny, int nz, int bx, int by, int bz, int nranks, const uint32_t *owner, ehyb_grid_decomp **out);
int ehyb_grid_decomp_info(const ehyb_grid_decomp *D, int64_t *nBricks, const int64_t **rowStarts, const int32_t **owner);
void ehyb_grid_decomp_free(ehyb_grid_decomp *D);
/* er_fill, exchange: as ehyb_mg_local_finish; chunkBricks bricks per builder call (0 = 64). */
int ehyb_mg_grid_build(const ehyb_grid_decomp *D, int rank, double er_fill, int exchange, int chunkBricks, ehyb_mg_local **out);
/* natural grid index (z*ny + y)*nx + x of every local row in permuted order (to fill x, check y) */
int ehyb_mg_local_natural_ids(const ehyb_mg_local *L, int64_t *out);
/* natural grid index of level-1 ids (e.g. of a halo list) */
int ehyb_grid_natural_ids(const ehyb_grid_decomp *D, int64_t count, const int64_t *level1, int64_t *out);
/* all rows of a rank at once, level-1 order and ids, partVec = brick of the row: the input of
 * the general path (ehyb_mg_local_build + ehyb_mg_local_finish) - small grids, parity tests */
int ehyb_grid_rows(const ehyb_grid_decomp *D, int rank, int64_t **rowPtr, int64_t **col, double **val, uint32_t **partVec);

/* pinned host memory for asynchronous host-vector products */
int ehyb_host_alloc_pinned(size_t bytes, void **out);
int ehyb_host_free_pinned(void *p);
/* launch geometry chosen by ehyb_upload */
int ehyb_session_info(const ehyb_handle *h, int *threads, int *ctasPerSM, int *grid, int64_t *smemBytes,
                      int *l2_persist);

/* ------------------------------------------------------------------------------------ */
/* synthetic matrices (BASELINE.json configs) and Matrix Market I/O    */
/* ------------------------------------------------------------------------------------ */

typedef enum ehyb_gen_kind {
    EHYB_GEN_LAPLACE2D = 1, /* 5-point, diag 4, off -1; dims nx, ny */
    EHYB_GEN_STENCIL27 = 2, /* 27-point, diag 26, off -1; dims nx, ny, nz */
    EHYB_GEN_ELASTICITY = 3 /* 3 dof/node, 27-point node stencil, dense 3x3 blocks */
} ehyb_gen_kind;

/* Lower-triangle "file entries" in the column-major order a .mtx of the matrix holds.
 * Arrays are malloc'd (free with ehyb_free_host).  0-based. */
int ehyb_gen_lower(ehyb_gen_kind kind, int nx, int ny, int nz, int *n, int64_t *count, int **li,
                   int **lj, double **lv);
/* The reader's expansion of a symmetric file into the matrixCOO the pipeline starts from
 * (solver_test.c:127-265): row-sorted COO with both triangles, rowIdx, numInRow, diag, maxCol,
 * zeroed numInRow2/partBoundary/reorderList.  y_golden (optional, n, zeroed by the callee) is
 * accumulated in file order like the reference's check vector; x may be NULL then. */
int ehyb_coo_from_lower(int n, int64_t count, const int *li, const int *lj, const double *lv,
                        matrixCOO *out, const double *x, double *y_golden);
/* General (unsymmetric) entries in file order (solver_test.c:31-126). */
int ehyb_coo_from_general(int n, int64_t count, const int *fi, const int *fj, const double *fv,
                          matrixCOO *out, const double *x, double *y_golden);
/* R-MAT (a,b,c,d)=(0.57,0.19,0.19,0.05), counter-based hash, duplicates summed, values
 * U(-1,1); general entries sorted by (row, col).  add_diagonal adds 4.0 on the diagonal. */
int ehyb_gen_rmat(int scale, int edge_factor, uint64_t seed, int add_diagonal, int *n,
                  int64_t *count, int **fi, int **fj, double **fv);
/* x of the reference driver: srand(i); x[i] = (rand()%200-100)/1000 (solver_test.c:228-232). */
void ehyb_x_reference(int n, double *x);
/* Reads a Matrix Market coordinate file (real|integer|pattern, general|symmetric) with the
 * reference reader's semantics, but with a buffered parser instead of one fscanf per line.
 * x_out (optional) receives the driver's x (malloc'd), y_out (optional, needs x_out) the
 * golden product accumulated in file order. */
int ehyb_read_mtx(const char *path, matrixCOO *out, int *symmetric, double **x_out, double **y_out);
int ehyb_write_mtx(const char *path, int n, int64_t count, const int *i, const int *j, const double *v,
                   int symmetric);
void ehyb_coo_free(matrixCOO *m);

#ifdef __cplusplus
}
#endif
#endif
