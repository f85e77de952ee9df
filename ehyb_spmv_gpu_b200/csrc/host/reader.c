/*
 * reader.c -- Matrix Market ingest with the reference reader's semantics
 * (solver_test.c:328-355 banner handling, :127-265 symmetric expansion, :31-126 general),
 * but parsing the entry lines from one buffered read instead of one fscanf per line
 * (the reference's largest host cost: 11.5 s for 29 M lines).
 */
#include <ctype.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.h"
#include "mmio.h"

static const char *skip_ws(const char *p, const char *end)
{
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    return p;
}

static const char *parse_int(const char *p, const char *end, int *out, int *ok)
{
    p = skip_ws(p, end);
    int neg = 0;
    long v = 0;
    if (p < end && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
    if (p >= end || !isdigit((unsigned char)*p)) { *ok = 0; return p; }
    while (p < end && isdigit((unsigned char)*p)) {
        v = v * 10 + (*p - '0');
        if (v > INT_MAX) { *ok = 0; return p; }
        ++p;
    }
    *out = (int)(neg ? -v : v);
    return p;
}

int ehyb_read_mtx(const char *path, matrixCOO *out, int *symmetric, double **x_out, double **y_out)
{
    if (!path || !out || !symmetric) return ehyb_fail(EHYB_ERR_ARG, "ehyb_read_mtx: NULL argument");
    FILE *f = fopen(path, "rb");
    if (!f) return ehyb_fail(EHYB_ERR_IO, "cannot open %s", path);
    MM_typecode tc;
    int M = 0, N = 0, nz = 0, rc = EHYB_OK;
    char *buf = NULL;
    int *ei = NULL, *ej = NULL;
    double *ev = NULL, *x = NULL, *y = NULL;
    if (mm_read_banner(f, &tc) != 0) { rc = ehyb_fail(EHYB_ERR_IO, "%s: could not process the Matrix Market banner", path); goto done; }
    if (!mm_is_matrix(tc) || !mm_is_sparse(tc) || mm_is_complex(tc)) {
        char *s = mm_typecode_to_str(tc);
        rc = ehyb_fail(EHYB_ERR_IO, "%s: unsupported Matrix Market type [%s]", path, s ? s : "?");
        free(s);
        goto done;
    }
    if (mm_read_mtx_crd_size(f, &M, &N, &nz) != 0 || M <= 0 || M != N || nz < 0) {
        rc = ehyb_fail(EHYB_ERR_IO, "%s: bad size line (%d x %d, %d entries); square matrices only", path, M, N, nz);
        goto done;
    }
    long pos = ftell(f);
    fseek(f, 0, SEEK_END);
    long size = ftell(f) - pos;
    fseek(f, pos, SEEK_SET);
    buf = (char *)malloc((size_t)size + 1);
    ei = (int *)malloc((size_t)(nz ? nz : 1) * sizeof(int));
    ej = (int *)malloc((size_t)(nz ? nz : 1) * sizeof(int));
    ev = (double *)malloc((size_t)(nz ? nz : 1) * sizeof(double));
    if (!buf || !ei || !ej || !ev) { rc = ehyb_fail(EHYB_ERR_NOMEM, "reader: out of memory"); goto done; }
    if (fread(buf, 1, (size_t)size, f) != (size_t)size) { rc = ehyb_fail(EHYB_ERR_IO, "%s: short read", path); goto done; }
    buf[size] = 0;
    const char *p = buf, *end = buf + size;
    const int pattern = mm_is_pattern(tc);
    for (int k = 0; k < nz; ++k) {
        int ok = 1, r = 0, c = 0;
        p = parse_int(p, end, &r, &ok);
        p = parse_int(p, end, &c, &ok);
        double v = 1.0;
        if (ok && !pattern) {
            p = skip_ws(p, end);
            char *q;
            v = strtod(p, &q);
            if (q == p) ok = 0;
            p = q;
        }
        if (!ok || r < 1 || r > M || c < 1 || c > N) { rc = ehyb_fail(EHYB_ERR_IO, "%s: bad entry line %d", path, k + 1); goto done; }
        ei[k] = r - 1; ej[k] = c - 1; ev[k] = v; /* 1-based -> 0-based, solver_test.c:98-99 */
    }
    *symmetric = mm_is_symmetric(tc) ? 1 : 0;
    if (x_out) {
        x = (double *)malloc((size_t)M * sizeof(double));
        if (!x) { rc = ehyb_fail(EHYB_ERR_NOMEM, "reader: out of memory"); goto done; }
        ehyb_x_reference(M, x);
    }
    if (y_out && x) {
        y = (double *)calloc((size_t)M, sizeof(double)); /* the reference mallocs and relies on fresh pages, B-10 */
        if (!y) { rc = ehyb_fail(EHYB_ERR_NOMEM, "reader: out of memory"); goto done; }
    }
    rc = *symmetric ? ehyb_coo_from_lower(M, nz, ei, ej, ev, out, x, y) : ehyb_coo_from_general(M, nz, ei, ej, ev, out, x, y);
    if (rc == EHYB_OK) {
        if (x_out) { *x_out = x; x = NULL; }
        if (y_out) { *y_out = y; y = NULL; }
    }
done:
    fclose(f);
    free(buf); free(ei); free(ej); free(ev); free(x); free(y);
    return rc;
}

int ehyb_write_mtx(const char *path, int n, int64_t count, const int *i, const int *j, const double *v, int symmetric)
{
    FILE *f = fopen(path, "w");
    if (!f) return ehyb_fail(EHYB_ERR_IO, "cannot create %s", path);
    MM_typecode tc;
    mm_initialize_typecode(&tc);
    mm_set_matrix(&tc); mm_set_coordinate(&tc); mm_set_real(&tc);
    if (symmetric) mm_set_symmetric(&tc); else mm_set_general(&tc);
    int bad = mm_write_banner(f, tc) || mm_write_mtx_crd_size(f, n, n, (int)count);
    for (int64_t k = 0; k < count && !bad; ++k) bad = fprintf(f, "%d %d %.17g\n", i[k] + 1, j[k] + 1, v[k]) <= 0;
    bad |= fclose(f) != 0;
    return bad ? ehyb_fail(EHYB_ERR_IO, "writing %s failed", path) : EHYB_OK;
}
