/*
 * cache.c -- binary cache of a finished pipeline.
 *
 * The reference re-runs reader -> mt-metis -> reorder -> COO2EHYB on every invocation
 * (solver_test.c:267-408; at config 2: 11.5 s of fscanf, 8.6 s of partitioning, 0.6 s of format
 * build).  The result of all that is a pure function of the .mtx file and
 * the partition parameters, so it is written once next to the file and mapped back in on the
 * next run: permutation, x, golden y, |A||x| (for the accuracy gate) and the tuned layout.
 *
 * File = header | int32 reorderList[n] | double x[n] | double yGolden[n] | double absAx[n] |
 *        layout (view scalars, then every array of struct ehyb_layout), all little-endian,
 *        every array padded to 8 bytes.  The header carries the identity of the source file
 *        (size + mtime) and the partition parameters; a 64-bit FNV-1a over the payload guards
 *        against truncation.  A cache that does not match is rejected (EHYB_ERR_IO), never
 *        silently used.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include "common.h"

#define EHYB_CACHE_MAGIC "EHYBB2\0\1" /* 8 bytes, the last one is the format version */

typedef struct {
    char magic[8];
    int64_t n, ncols, nnz;
    int32_t nParts, W, ctasPerPart, nSlices;
    int64_t blobBytes, nOverflow, cacheTotal;
    int32_t cacheMax, haloInOverflow;
    int64_t nnzEll, nnzRemInSlice, nnzOverflow, padEll, padRem, nLongRows, algBytes, formatBytes;
    int64_t sourceSize, sourceMtimeSec, sourceMtimeNsec;
    int32_t hasVectors, symmetric;
    uint64_t checksum; /* FNV-1a 64 over everything after the header */
    int64_t reserved[4];
} cache_header;

/* Options the layout was built with that the header's partition parameters do not tell (er_fill,
 * cache_cap, the partitioner and its pieces, ...): the caller folds them into one 64-bit tag
 * (ehyb_cache_set_options_tag); a cache written under another tag is rejected like one with other
 * partition parameters.  0 (the default) on both sides = no options to tell apart. */
static uint64_t g_optionsTag = 0;
void ehyb_cache_set_options_tag(uint64_t tag) { g_optionsTag = tag; }

/* layout.c */
int ehyb_layout_export_arrays(const ehyb_layout *L, const void **arrays, int64_t *bytes, int max);
int ehyb_layout_import(const ehyb_layout_view *scalars, void *const *arrays, ehyb_layout **out);
#define EHYB_LAYOUT_ARRAYS 11

static uint64_t fnv1a(uint64_t h, const void *p, size_t len)
{
    /* 8 bytes at a time: this is an integrity check of ~600 MB files, not a cryptographic hash */
    const uint64_t *w = (const uint64_t *)p;
    size_t nw = len / 8;
    for (size_t i = 0; i < nw; ++i) { h ^= w[i]; h *= 0x100000001b3ULL; }
    const unsigned char *b = (const unsigned char *)p + nw * 8;
    for (size_t i = 0; i < len % 8; ++i) { h ^= b[i]; h *= 0x100000001b3ULL; }
    return h;
}

static int source_identity(const char *source_path, cache_header *h)
{
    h->sourceSize = h->sourceMtimeSec = h->sourceMtimeNsec = 0;
    if (!source_path || !source_path[0]) return EHYB_OK;
    struct stat st;
    if (stat(source_path, &st) != 0) return ehyb_fail(EHYB_ERR_IO, "cache: cannot stat %s", source_path);
    h->sourceSize = (int64_t)st.st_size;
    h->sourceMtimeSec = (int64_t)st.st_mtim.tv_sec;
    h->sourceMtimeNsec = (int64_t)st.st_mtim.tv_nsec;
    return EHYB_OK;
}

static int put(FILE *f, const void *p, int64_t bytes, uint64_t *sum)
{
    static const char zero[8] = {0};
    if (bytes > 0 && fwrite(p, 1, (size_t)bytes, f) != (size_t)bytes) return -1;
    *sum = fnv1a(*sum, p, (size_t)bytes);
    const size_t pad = (size_t)((8 - bytes % 8) % 8);
    if (pad && fwrite(zero, 1, pad, f) != pad) return -1;
    return 0;
}

static void *get(FILE *f, int64_t bytes, uint64_t *sum, int *err)
{
    void *p = malloc((size_t)(bytes > 0 ? bytes : 8));
    if (!p) { *err = EHYB_ERR_NOMEM; return NULL; }
    char pad[8];
    if ((bytes > 0 && fread(p, 1, (size_t)bytes, f) != (size_t)bytes) ||
        ((8 - bytes % 8) % 8 && fread(pad, 1, (size_t)((8 - bytes % 8) % 8), f) != (size_t)((8 - bytes % 8) % 8))) {
        free(p);
        *err = EHYB_ERR_IO;
        return NULL;
    }
    *sum = fnv1a(*sum, p, (size_t)bytes);
    return p;
}

int ehyb_cache_save(const char *path, const char *source_path, const ehyb_layout *L, int symmetric, const int *reorderList,
                    const double *x, const double *y_golden, const double *absAx)
{
    if (!path || !L) return ehyb_fail(EHYB_ERR_ARG, "ehyb_cache_save: NULL argument");
    ehyb_layout_view v;
    int rc = ehyb_layout_get(L, &v);
    if (rc) return rc;
    cache_header h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, EHYB_CACHE_MAGIC, 8);
    h.n = v.n; h.ncols = v.ncols; h.nnz = v.nnz; h.nParts = v.nParts; h.W = v.W; h.ctasPerPart = v.ctasPerPart; h.nSlices = v.nSlices;
    h.blobBytes = v.blobBytes; h.nOverflow = v.nOverflow; h.cacheTotal = v.cacheTotal; h.cacheMax = v.cacheMax;
    h.haloInOverflow = v.haloInOverflow;
    h.nnzEll = v.nnzEll; h.nnzRemInSlice = v.nnzRemInSlice; h.nnzOverflow = v.nnzOverflow; h.padEll = v.padEll; h.padRem = v.padRem;
    h.nLongRows = v.nLongRows; h.algBytes = v.algBytes; h.formatBytes = v.formatBytes;
    h.hasVectors = reorderList && x && y_golden && absAx;
    h.symmetric = symmetric;
    h.reserved[0] = (int64_t)g_optionsTag;
    if ((rc = source_identity(source_path, &h))) return rc;
    char tmp[4096];
    snprintf(tmp, sizeof tmp, "%s.tmp%ld", path, (long)getpid());
    FILE *f = fopen(tmp, "wb");
    if (!f) return ehyb_fail(EHYB_ERR_IO, "cache: cannot create %s", tmp);
    uint64_t sum = 0xcbf29ce484222325ULL;
    int bad = fwrite(&h, sizeof h, 1, f) != 1;
    if (!bad && h.hasVectors) {
        bad = put(f, reorderList, 4 * v.n, &sum) || put(f, x, 8 * v.n, &sum) || put(f, y_golden, 8 * v.n, &sum) || put(f, absAx, 8 * v.n, &sum);
    }
    const void *arr[EHYB_LAYOUT_ARRAYS];
    int64_t bytes[EHYB_LAYOUT_ARRAYS];
    const int na = ehyb_layout_export_arrays(L, arr, bytes, EHYB_LAYOUT_ARRAYS);
    for (int i = 0; i < na && !bad; ++i) bad = put(f, arr[i], bytes[i], &sum);
    if (!bad) {
        h.checksum = sum;
        bad = fseek(f, 0, SEEK_SET) != 0 || fwrite(&h, sizeof h, 1, f) != 1;
    }
    bad = fclose(f) != 0 || bad;
    if (bad || rename(tmp, path) != 0) {
        remove(tmp);
        return ehyb_fail(EHYB_ERR_IO, "cache: writing %s failed", path);
    }
    return EHYB_OK;
}

int ehyb_cache_load(const char *path, const char *source_path, const ehyb_plan_t *plan, ehyb_layout **L_out, int *n_out,
                    int *symmetric, int **reorderList, double **x, double **y_golden, double **absAx)
{
    if (!path || !L_out) return ehyb_fail(EHYB_ERR_ARG, "ehyb_cache_load: NULL argument");
    FILE *f = fopen(path, "rb");
    if (!f) return ehyb_fail(EHYB_ERR_IO, "cache: %s not found", path);
    cache_header h, now;
    int rc = EHYB_OK, err = 0;
    void *arr[EHYB_LAYOUT_ARRAYS] = {0};
    int *rl = NULL;
    double *vx = NULL, *vy = NULL, *va = NULL;
    uint64_t sum = 0xcbf29ce484222325ULL;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, EHYB_CACHE_MAGIC, 8) != 0) { rc = ehyb_fail(EHYB_ERR_IO, "cache: %s is not an EHYB cache of this version", path); goto done; }
    memset(&now, 0, sizeof now);
    if ((rc = source_identity(source_path, &now))) goto done;
    if (source_path && source_path[0] && (now.sourceSize != h.sourceSize || now.sourceMtimeSec != h.sourceMtimeSec || now.sourceMtimeNsec != h.sourceMtimeNsec)) {
        rc = ehyb_fail(EHYB_ERR_IO, "cache: %s was built from another version of %s", path, source_path);
        goto done;
    }
    if (plan && (plan->nParts != h.nParts || plan->W != h.W || (plan->ctasPerPart > 0 ? plan->ctasPerPart : 1) != h.ctasPerPart)) {
        rc = ehyb_fail(EHYB_ERR_IO, "cache: %s holds P=%d W=%d K=%d, wanted P=%d W=%d K=%d", path, h.nParts, h.W, h.ctasPerPart, plan->nParts, plan->W, plan->ctasPerPart);
        goto done;
    }
    if ((uint64_t)h.reserved[0] != g_optionsTag) {
        rc = ehyb_fail(EHYB_ERR_IO, "cache: %s was built with other layout / partitioner options (tag %llx, this run %llx)", path,
                       (unsigned long long)h.reserved[0], (unsigned long long)g_optionsTag);
        goto done;
    }
    if (h.n <= 0 || h.n > 0x7fffffff || h.nParts <= 0 || h.nSlices < 0 || h.blobBytes < 0 || h.nOverflow < 0 || h.cacheTotal < 0) { rc = ehyb_fail(EHYB_ERR_IO, "cache: %s has an inconsistent header", path); goto done; }
    if (h.hasVectors) {
        rl = (int *)get(f, 4 * h.n, &sum, &err);
        if (!err) vx = (double *)get(f, 8 * h.n, &sum, &err);
        if (!err) vy = (double *)get(f, 8 * h.n, &sum, &err);
        if (!err) va = (double *)get(f, 8 * h.n, &sum, &err);
    }
    {
        const int64_t bytes[EHYB_LAYOUT_ARRAYS] = {
            (int64_t)sizeof(ehyb_part_desc) * h.nParts, (int64_t)sizeof(ehyb_slice_desc) * h.nSlices, h.blobBytes,
            4 * h.nOverflow, 4 * h.nOverflow, 8 * h.nOverflow, 4 * h.cacheTotal,
            4 * h.n, 4 * h.n, 4 * h.n, 8 * (h.n + 1)};
        for (int i = 0; i < EHYB_LAYOUT_ARRAYS && !err; ++i) arr[i] = get(f, bytes[i], &sum, &err);
    }
    if (err) { rc = ehyb_fail(err, err == EHYB_ERR_NOMEM ? "cache: out of memory" : "cache: %s is truncated", path); goto done; }
    if (sum != h.checksum) { rc = ehyb_fail(EHYB_ERR_IO, "cache: %s fails its checksum", path); goto done; }
    {
        ehyb_layout_view v;
        memset(&v, 0, sizeof v);
        v.n = h.n; v.ncols = h.ncols; v.nnz = h.nnz; v.nParts = h.nParts; v.W = h.W; v.ctasPerPart = h.ctasPerPart; v.nSlices = h.nSlices;
        v.blobBytes = h.blobBytes; v.nOverflow = h.nOverflow; v.cacheTotal = h.cacheTotal; v.cacheMax = h.cacheMax; v.haloInOverflow = h.haloInOverflow;
        v.nnzEll = h.nnzEll; v.nnzRemInSlice = h.nnzRemInSlice; v.nnzOverflow = h.nnzOverflow; v.padEll = h.padEll; v.padRem = h.padRem;
        v.nLongRows = h.nLongRows; v.algBytes = h.algBytes; v.formatBytes = h.formatBytes;
        rc = ehyb_layout_import(&v, arr, L_out); /* takes ownership of the arrays on success */
        if (rc == EHYB_OK) memset(arr, 0, sizeof arr);
    }
    if (rc == EHYB_OK) {
        if (n_out) *n_out = (int)h.n;
        if (symmetric) *symmetric = h.symmetric;
        if (reorderList) { *reorderList = rl; rl = NULL; }
        if (x) { *x = vx; vx = NULL; }
        if (y_golden) { *y_golden = vy; vy = NULL; }
        if (absAx) { *absAx = va; va = NULL; }
    }
done:
    fclose(f);
    for (int i = 0; i < EHYB_LAYOUT_ARRAYS; ++i) free(arr[i]);
    free(rl); free(vx); free(vy); free(va);
    return rc;
}

/* the layout alone (no source identity, no vectors) */
int ehyb_layout_save(const ehyb_layout *L, const char *path) { return ehyb_cache_save(path, NULL, L, 1, NULL, NULL, NULL, NULL); }
int ehyb_layout_load(const char *path, ehyb_layout **out) { return ehyb_cache_load(path, NULL, NULL, out, NULL, NULL, NULL, NULL, NULL, NULL); }
