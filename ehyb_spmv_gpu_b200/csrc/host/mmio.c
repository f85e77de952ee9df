/*
 * mmio.c -- Matrix Market banner / size-line I/O (interface: include/mmio.h).
 * Written from scratch against the published Matrix Market exchange format; behaviour
 * (return codes, accepted spellings, case-insensitivity) follows the NIST library the
 * reference ships (reference mmio.c:96-217, :455-511).
 */
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include "mmio.h"

static void lower(char *s)
{
    for (; *s; ++s) *s = (char)tolower((unsigned char)*s);
}

int mm_is_valid(MM_typecode t)
{
    if (!mm_is_matrix(t)) return 0;
    if (mm_is_dense(t) && mm_is_pattern(t)) return 0;
    if (mm_is_real(t) && mm_is_hermitian(t)) return 0;
    if (mm_is_pattern(t) && (mm_is_hermitian(t) || mm_is_skew(t))) return 0;
    return 1;
}

int mm_read_banner(FILE *f, MM_typecode *matcode)
{
    char line[MM_MAX_LINE_LENGTH];
    char w[5][MM_MAX_TOKEN_LENGTH + 1];
    mm_clear_typecode(matcode);
    if (!fgets(line, sizeof line, f)) return MM_PREMATURE_EOF;
    if (sscanf(line, "%64s %64s %64s %64s %64s", w[0], w[1], w[2], w[3], w[4]) != 5) return MM_PREMATURE_EOF;
    for (int i = 1; i < 5; ++i) lower(w[i]);
    if (strncmp(w[0], MatrixMarketBanner, strlen(MatrixMarketBanner)) != 0) return MM_NO_HEADER;
    if (strcmp(w[1], MM_MTX_STR) != 0) return MM_UNSUPPORTED_TYPE;
    mm_set_matrix(matcode);

    if (!strcmp(w[2], MM_SPARSE_STR)) mm_set_sparse(matcode);
    else if (!strcmp(w[2], MM_DENSE_STR)) mm_set_dense(matcode);
    else return MM_UNSUPPORTED_TYPE;

    if (!strcmp(w[3], MM_REAL_STR)) mm_set_real(matcode);
    else if (!strcmp(w[3], MM_COMPLEX_STR)) mm_set_complex(matcode);
    else if (!strcmp(w[3], MM_PATTERN_STR)) mm_set_pattern(matcode);
    else if (!strcmp(w[3], MM_INT_STR)) mm_set_integer(matcode);
    else return MM_UNSUPPORTED_TYPE;

    if (!strcmp(w[4], MM_GENERAL_STR)) mm_set_general(matcode);
    else if (!strcmp(w[4], MM_SYMM_STR)) mm_set_symmetric(matcode);
    else if (!strcmp(w[4], MM_HERM_STR)) mm_set_hermitian(matcode);
    else if (!strcmp(w[4], MM_SKEW_STR)) mm_set_skew(matcode);
    else return MM_UNSUPPORTED_TYPE;
    return 0;
}

/* first line that is not a % comment and not blank */
static int next_data_line(FILE *f, char *line, int cap)
{
    for (;;) {
        if (!fgets(line, cap, f)) return MM_PREMATURE_EOF;
        const char *p = line;
        while (*p == ' ' || *p == '\t') ++p;
        if (*p == '%' || *p == '\n' || *p == '\r' || *p == 0) continue;
        return 0;
    }
}

int mm_read_mtx_crd_size(FILE *f, int *M, int *N, int *nz)
{
    char line[MM_MAX_LINE_LENGTH];
    *M = *N = *nz = 0;
    for (;;) {
        int rc = next_data_line(f, line, sizeof line);
        if (rc) return rc;
        if (sscanf(line, "%d %d %d", M, N, nz) == 3) return 0;
    }
}

int mm_read_mtx_array_size(FILE *f, int *M, int *N)
{
    char line[MM_MAX_LINE_LENGTH];
    *M = *N = 0;
    for (;;) {
        int rc = next_data_line(f, line, sizeof line);
        if (rc) return rc;
        if (sscanf(line, "%d %d", M, N) == 2) return 0;
    }
}

char *mm_typecode_to_str(MM_typecode t)
{
    const char *w1 = mm_is_matrix(t) ? MM_MTX_STR : NULL;
    const char *w2 = mm_is_sparse(t) ? MM_SPARSE_STR : mm_is_dense(t) ? MM_DENSE_STR : NULL;
    const char *w3 = mm_is_real(t) ? MM_REAL_STR : mm_is_complex(t) ? MM_COMPLEX_STR
                   : mm_is_pattern(t) ? MM_PATTERN_STR : mm_is_integer(t) ? MM_INT_STR : NULL;
    const char *w4 = mm_is_general(t) ? MM_GENERAL_STR : mm_is_symmetric(t) ? MM_SYMM_STR
                   : mm_is_hermitian(t) ? MM_HERM_STR : mm_is_skew(t) ? MM_SKEW_STR : NULL;
    if (!w1 || !w2 || !w3 || !w4) return NULL;
    char *s = (char *)malloc(MM_MAX_LINE_LENGTH);
    if (s) snprintf(s, MM_MAX_LINE_LENGTH, "%s %s %s %s", w1, w2, w3, w4);
    return s;
}

int mm_write_banner(FILE *f, MM_typecode t)
{
    char *s = mm_typecode_to_str(t);
    int ok = s && fprintf(f, "%s %s\n", MatrixMarketBanner, s) > 0;
    free(s);
    return ok ? 0 : MM_COULD_NOT_WRITE_FILE;
}

int mm_write_mtx_crd_size(FILE *f, int M, int N, int nz)
{
    return fprintf(f, "%d %d %d\n", M, N, nz) > 0 ? 0 : MM_COULD_NOT_WRITE_FILE;
}

int mm_write_mtx_array_size(FILE *f, int M, int N)
{
    return fprintf(f, "%d %d\n", M, N) > 0 ? 0 : MM_COULD_NOT_WRITE_FILE;
}

int mm_read_mtx_crd_entry(FILE *f, int *I, int *J, double *re, double *im, MM_typecode t)
{
    if (mm_is_complex(t)) return fscanf(f, "%d %d %lg %lg", I, J, re, im) == 4 ? 0 : MM_PREMATURE_EOF;
    if (mm_is_real(t) || mm_is_integer(t)) return fscanf(f, "%d %d %lg", I, J, re) == 3 ? 0 : MM_PREMATURE_EOF;
    if (mm_is_pattern(t)) return fscanf(f, "%d %d", I, J) == 2 ? 0 : MM_PREMATURE_EOF;
    return MM_UNSUPPORTED_TYPE;
}

int mm_read_mtx_crd_data(FILE *f, int M, int N, int nz, int I[], int J[], double val[], MM_typecode t)
{
    (void)M; (void)N;
    for (int k = 0; k < nz; ++k) {
        double re = 1.0, im = 0.0;
        int rc = mm_read_mtx_crd_entry(f, &I[k], &J[k], &re, &im, t);
        if (rc) return rc;
        if (mm_is_complex(t)) { val[2 * k] = re; val[2 * k + 1] = im; }
        else if (!mm_is_pattern(t)) val[k] = re;
    }
    return 0;
}

int mm_write_mtx_crd(char fname[], int M, int N, int nz, int I[], int J[], double val[], MM_typecode t)
{
    FILE *f = strcmp(fname, "stdout") == 0 ? stdout : fopen(fname, "w");
    if (!f) return MM_COULD_NOT_WRITE_FILE;
    int rc = mm_write_banner(f, t);
    if (!rc) rc = mm_write_mtx_crd_size(f, M, N, nz);
    for (int k = 0; k < nz && !rc; ++k) {
        int ok;
        if (mm_is_pattern(t)) ok = fprintf(f, "%d %d\n", I[k], J[k]);
        else if (mm_is_complex(t)) ok = fprintf(f, "%d %d %20.16g %20.16g\n", I[k], J[k], val[2 * k], val[2 * k + 1]);
        else ok = fprintf(f, "%d %d %20.16g\n", I[k], J[k], val[k]);
        if (ok <= 0) rc = MM_COULD_NOT_WRITE_FILE;
    }
    if (f != stdout) fclose(f);
    return rc;
}

int mm_read_unsymmetric_sparse(const char *fname, int *M_, int *N_, int *nz_, double **val_, int **I_, int **J_)
{
    FILE *f = fopen(fname, "r");
    if (!f) return -1;
    MM_typecode t;
    int M, N, nz;
    if (mm_read_banner(f, &t) || !(mm_is_real(t) && mm_is_matrix(t) && mm_is_sparse(t)) ||
        mm_read_mtx_crd_size(f, &M, &N, &nz)) {
        fclose(f);
        return -1;
    }
    int *I = (int *)malloc((size_t)(nz ? nz : 1) * sizeof(int)), *J = (int *)malloc((size_t)(nz ? nz : 1) * sizeof(int));
    double *v = (double *)malloc((size_t)(nz ? nz : 1) * sizeof(double));
    int bad = !I || !J || !v;
    for (int k = 0; k < nz && !bad; ++k) {
        bad = fscanf(f, "%d %d %lg", &I[k], &J[k], &v[k]) != 3;
        I[k]--; J[k]--;
    }
    fclose(f);
    if (bad) { free(I); free(J); free(v); return -1; }
    *M_ = M; *N_ = N; *nz_ = nz; *val_ = v; *I_ = I; *J_ = J;
    return 0;
}
