/*
 * mmio.h -- Matrix Market banner / size-line I/O.
 *
 * API-compatible with the NIST "Matrix Market I/O library for ANSI C" interface the reference
 * ships (reference mmio.h:18-129): same function names, typecode letters and error codes, so
 * that solver_test.c and user code written against it keep compiling.  The implementation
 * (host/mmio.c) is written from scratch; the driver only needs the banner, the size line
 * and the typecode predicates (solver_test.c:34,131,333-348) - entry lines are parsed by
 * host/reader.c.
 */
#ifndef MM_IO_H
#define MM_IO_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MM_MAX_LINE_LENGTH 1025
#define MM_MAX_TOKEN_LENGTH 64
#define MatrixMarketBanner "%%MatrixMarket"

/* [0] object 'M'; [1] format 'C'oordinate|'A'rray; [2] field 'R'eal|'C'omplex|'P'attern|
 * 'I'nteger; [3] symmetry 'G'eneral|'S'ymmetric|'H'ermitian|s'K'ew */
typedef char MM_typecode[4];

/* error codes */
#define MM_COULD_NOT_READ_FILE 11
#define MM_PREMATURE_EOF 12
#define MM_NOT_MTX 13
#define MM_NO_HEADER 14
#define MM_UNSUPPORTED_TYPE 15
#define MM_LINE_TOO_LONG 16
#define MM_COULD_NOT_WRITE_FILE 17

/* banner words */
#define MM_MTX_STR "matrix"
#define MM_ARRAY_STR "array"
#define MM_DENSE_STR "array"
#define MM_COORDINATE_STR "coordinate"
#define MM_SPARSE_STR "coordinate"
#define MM_COMPLEX_STR "complex"
#define MM_REAL_STR "real"
#define MM_INT_STR "integer"
#define MM_PATTERN_STR "pattern"
#define MM_GENERAL_STR "general"
#define MM_SYMM_STR "symmetric"
#define MM_HERM_STR "hermitian"
#define MM_SKEW_STR "skew-symmetric"

/* predicates */
#define mm_is_matrix(t) ((t)[0] == 'M')
#define mm_is_sparse(t) ((t)[1] == 'C')
#define mm_is_coordinate(t) ((t)[1] == 'C')
#define mm_is_dense(t) ((t)[1] == 'A')
#define mm_is_array(t) ((t)[1] == 'A')
#define mm_is_complex(t) ((t)[2] == 'C')
#define mm_is_real(t) ((t)[2] == 'R')
#define mm_is_pattern(t) ((t)[2] == 'P')
#define mm_is_integer(t) ((t)[2] == 'I')
#define mm_is_symmetric(t) ((t)[3] == 'S')
#define mm_is_general(t) ((t)[3] == 'G')
#define mm_is_skew(t) ((t)[3] == 'K')
#define mm_is_hermitian(t) ((t)[3] == 'H')

/* setters (take a pointer to the typecode) */
#define mm_set_matrix(t) ((*(t))[0] = 'M')
#define mm_set_coordinate(t) ((*(t))[1] = 'C')
#define mm_set_array(t) ((*(t))[1] = 'A')
#define mm_set_dense(t) mm_set_array(t)
#define mm_set_sparse(t) mm_set_coordinate(t)
#define mm_set_complex(t) ((*(t))[2] = 'C')
#define mm_set_real(t) ((*(t))[2] = 'R')
#define mm_set_pattern(t) ((*(t))[2] = 'P')
#define mm_set_integer(t) ((*(t))[2] = 'I')
#define mm_set_symmetric(t) ((*(t))[3] = 'S')
#define mm_set_general(t) ((*(t))[3] = 'G')
#define mm_set_skew(t) ((*(t))[3] = 'K')
#define mm_set_hermitian(t) ((*(t))[3] = 'H')
#define mm_clear_typecode(t) ((*(t))[0] = (*(t))[1] = (*(t))[2] = ' ', (*(t))[3] = 'G')
#define mm_initialize_typecode(t) mm_clear_typecode(t)

int mm_read_banner(FILE *f, MM_typecode *matcode);
int mm_read_mtx_crd_size(FILE *f, int *M, int *N, int *nz);
int mm_read_mtx_array_size(FILE *f, int *M, int *N);
int mm_write_banner(FILE *f, MM_typecode matcode);
int mm_write_mtx_crd_size(FILE *f, int M, int N, int nz);
int mm_write_mtx_array_size(FILE *f, int M, int N);
int mm_is_valid(MM_typecode matcode);
/* Returns a malloc'd string "matrix coordinate real general" (caller frees). */
char *mm_typecode_to_str(MM_typecode matcode);

int mm_read_mtx_crd_entry(FILE *f, int *I, int *J, double *real, double *img, MM_typecode matcode);
int mm_read_mtx_crd_data(FILE *f, int M, int N, int nz, int I[], int J[], double val[], MM_typecode matcode);
int mm_write_mtx_crd(char fname[], int M, int N, int nz, int I[], int J[], double val[], MM_typecode matcode);
int mm_read_unsymmetric_sparse(const char *fname, int *M_, int *N_, int *nz_, double **val_, int **I_, int **J_);

#ifdef __cplusplus
}
#endif
#endif
