/*
 * ehyb_oracle.c -- CPU restatement of the reference EHYB SpMV path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker
 * (or as the reported CPU baseline), never as the thing measured or shipped.
 *
 * Parity status: PINNED.  Every function below is checked byte-for-byte
 * against the unmodified reference sources compiled in place
 * (oracle/_ref/libehyb_ref.so, see oracle/Makefile and
 * tests/test_oracle_pinned.py) and against the reference-derived
 * known-answer values of SURVEY.md Appendix D (tests/golden/).
 *
 * Each function cites the reference file:line it follows (paths are
 * relative to the reference checkout).  The code is sequential and written
 * for clarity, not speed; the only parallel routine is the CSR baseline.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_WARP 32
#define ORC_LONG 512 /* kernel.h:26 threadLongVec */

/* ------------------------------------------------------------------ */
/* utilities                                                           */
/* ------------------------------------------------------------------ */

/* FNV-1a 64 over raw bytes: the hash SURVEY.md Appendix D uses. */
uint64_t orc_fnv1a(const void *p, size_t nbytes)
{
    const unsigned char *b = (const unsigned char *)p;
    uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < nbytes; ++i) {
        h ^= b[i];
        h *= 0x100000001b3ULL;
    }
    return h;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* x generator of the driver: solver_test.c:228-232 (sym), :89-93 (unsym). */
void orc_x_reference(int n, double *x)
{
    for (int i = 0; i < n; ++i) {
        srand(i);
        x[i] = (double)(rand() % 200 - 100) / 1000;
    }
}

/* ------------------------------------------------------------------ */
/* reader semantics                                                    */
/* ------------------------------------------------------------------ */

/*
 * Symmetric reader: solver_test.c:127-265.  Input = the lower-triangle file
 * entries in file order (0-based).  Output = row-sorted COO with both
 * triangles, rowIdx, numInRow, diag, maxCol, and the golden y accumulated in
 * file order (y must come in zeroed: the reference relies on fresh pages,
 * SURVEY B-10).  totalNum = 2*lower - n (B-12: every diagonal present).
 * Returns totalNum.
 */
int orc_read_sym(int n, int lowerNum, const int *li, const int *lj, const double *lv,
                 int *I, int *J, double *V, int *rowIdx, int *numInRow, double *diag,
                 int *maxCol_out, const double *x, double *y)
{
    int totalNum = lowerNum * 2 - n;                      /* :135 */
    memset(numInRow, 0, sizeof(int) * (size_t)n);
    for (int i = 0; i < lowerNum; ++i) {                  /* :196-206 */
        numInRow[li[i]] += 1;
        if (li[i] != lj[i]) numInRow[lj[i]] += 1;
    }
    int maxCol = 0;
    rowIdx[0] = 0;
    for (int i = 1; i <= n; ++i) {                        /* :214-222 */
        if (numInRow[i - 1] > maxCol) maxCol = numInRow[i - 1];
        rowIdx[i] = rowIdx[i - 1] + numInRow[i - 1];
        numInRow[i - 1] = 0;
    }
    *maxCol_out = maxCol;
    for (int i = 0; i < lowerNum; ++i) {                  /* :235-260 */
        int tI = li[i], tJ = lj[i];
        double tV = lv[i];
        int index1 = rowIdx[tI] + numInRow[tI];
        int index2 = rowIdx[tJ] + numInRow[tJ];
        numInRow[tI] += 1;
        I[index1] = tI; J[index1] = tJ; V[index1] = tV;
        if (y) y[tI] += tV * x[tJ];
        if (tI != tJ) {
            numInRow[tJ] += 1;
            I[index2] = tJ; J[index2] = tI; V[index2] = tV;
            if (y) y[tJ] += tV * x[tI];
        } else {
            diag[tI] = tV;
        }
    }
    return totalNum;
}

/*
 * Unsymmetric reader: solver_test.c:31-126.  Entries stay in FILE order (the
 * COO is not row-sorted); rowIdx is only the prefix sum of the row counts.
 */
void orc_read_unsym(int n, int totalNum, const int *fi, const int *fj, const double *fv,
                    int *I, int *J, double *V, int *rowIdx, int *numInRow,
                    int *maxCol_out, const double *x, double *y)
{
    memset(numInRow, 0, sizeof(int) * (size_t)n);
    for (int i = 0; i < totalNum; ++i) {                  /* :96-103 */
        J[i] = fj[i]; I[i] = fi[i]; V[i] = fv[i];
        numInRow[fi[i]] += 1;
        if (y) y[fi[i]] += fv[i] * x[fj[i]];
    }
    int maxCol = 0;
    rowIdx[0] = 0;
    for (int i = 1; i <= n; ++i) {                        /* :111-121 */
        if (numInRow[i - 1] > maxCol) maxCol = numInRow[i - 1];
        rowIdx[i] = rowIdx[i - 1] + numInRow[i - 1];
    }
    *maxCol_out = maxCol;
}

/*
 * Partition-parameter heuristic: solver_test.c:158-182 (sym) / :53-77 (unsym),
 * with the constants of kernel.h:20-25 (smSize 82, smSize2 80,
 * maxSharedMem 93*1024, threadELL 1024).  The int16_t wrap of
 * vectorCacheSize (SURVEY B-7) is reproduced: gcc/x86 converts the
 * out-of-range double through a 32-bit int and truncates.
 * kpp_out = 0 means "left uninitialised by the reference" (B-13).
 */
static int16_t orc_to_i16(double d) { return (int16_t)(int32_t)d; }

void orc_heuristic_ref(int n, int symmetric, int *nParts_out, int *W_out, int *kpp_out)
{
    const int smSize = 82, smSize2 = 80, threadELL = 1024;
    const size_t maxSharedMem = 93 * 1024;
    int partFactor = 1, kernelPerPart = 1, nParts;
    int16_t W = orc_to_i16(ceil(((double)n) / (partFactor * smSize * threadELL)) * threadELL);
    /* int16 -> int -> size_t promotion in the comparison, as in the reference */
    if ((size_t)(long)W < maxSharedMem / (2 * sizeof(double))) {
        int kArray[4] = {8, 5, 4, 2};
        int kIdx = 0;
        kernelPerPart = kArray[kIdx];
        W = orc_to_i16(kernelPerPart * ceil(((double)n) / (smSize2 * threadELL)) * threadELL);
        kIdx++;
        while ((size_t)(long)W * sizeof(double) > maxSharedMem && kIdx < 4) {
            kernelPerPart = kArray[kIdx];
            W = orc_to_i16(kernelPerPart * ceil(((double)n) / (smSize2 * threadELL)) * threadELL);
            kIdx++;
        }
        nParts = (symmetric ? smSize2 : smSize) / kernelPerPart;   /* :173 vs :68 */
        *kpp_out = kernelPerPart;
    } else {
        while ((size_t)(long)W * sizeof(double) > maxSharedMem) {
            partFactor += 1;
            W = orc_to_i16(ceil(((double)n) / (partFactor * smSize * threadELL)) * threadELL);
        }
        nParts = partFactor * smSize;
        *kpp_out = 0;
    }
    *nParts_out = nParts;
    *W_out = (int)(uint16_t)W; /* stored into uint16_t matrixCOO.vectorCacheSize, spmv.h:22 */
}

/* ------------------------------------------------------------------ */
/* graph handed to mt-metis                                            */
/* ------------------------------------------------------------------ */

/* Symmetric path: reordering.c:239-264.  xadj = rowIdx, adjncy = J, both as
 * uint32, diagonal self-loops included (B-20). */
void orc_graph_sym(int n, int totalNum, const int *rowIdx, const int *J,
                   uint32_t *xadj, uint32_t *adjncy)
{
    for (int i = 0; i < totalNum; ++i) adjncy[i] = (uint32_t)J[i];
    for (int i = 0; i <= n; ++i) xadj[i] = (uint32_t)rowIdx[i];
}

/* Unsymmetric path: reordering.c:50-89.  Pattern of A + A^T, built by
 * scattering every entry (i,j) into row i and, when i != j, also into row j;
 * duplicates stay when both (i,j) and (j,i) exist (B-20).  expandNumInRow is
 * zeroed here (the reference relies on fresh pages, B-11).  adjncy must hold
 * 2*totalNum entries.  Returns the number of adjacency entries. */
uint32_t orc_graph_unsym(int n, int totalNum, const int *I, const int *J,
                         uint32_t *xadj, uint32_t *adjncy)
{
    uint32_t *cnt = (uint32_t *)calloc((size_t)n + 1, sizeof(uint32_t));
    for (int i = 0; i < totalNum; ++i) {                  /* :56-64 */
        cnt[I[i]] += 1;
        if (I[i] != J[i]) cnt[J[i]] += 1;
    }
    xadj[0] = 0;
    for (int i = 1; i <= n; ++i) {                        /* :65-69 */
        xadj[i] = xadj[i - 1] + cnt[i - 1];
        cnt[i - 1] = 0;
    }
    for (int i = 0; i < totalNum; ++i) {                  /* :71-89 */
        int tI = I[i], tJ = J[i];
        uint32_t index1 = xadj[tI] + cnt[tI];
        uint32_t index2 = xadj[tJ] + cnt[tJ];
        cnt[tI] += 1;
        adjncy[index1] = (uint32_t)tJ;
        if (tI != tJ) {
            cnt[tJ] += 1;
            adjncy[index2] = (uint32_t)tI;
        }
    }
    free(cnt);
    return xadj[n];
}

/* ------------------------------------------------------------------ */
/* reorder (everything after the mt-metis call)                        */
/* ------------------------------------------------------------------ */

typedef struct { unsigned int idx; unsigned int nonzeros; } orc_rowS; /* Partition.h:12-15 */

/* Partition.h:17-24: descending by nonzeros, 0 on ties.  The tie order is
 * made explicit (ascending idx == ascending position, because every caller
 * fills the array in ascending idx order): this is what glibc's merge-sort
 * qsort produces with the reference comparator. */
static int orc_rowS_cmp(const void *A, const void *B)
{
    const orc_rowS *a = (const orc_rowS *)A, *b = (const orc_rowS *)B;
    if (a->nonzeros > b->nonzeros) return -1;
    if (a->nonzeros < b->nonzeros) return 1;
    if (a->idx < b->idx) return -1;
    if (a->idx > b->idx) return 1;
    return 0;
}

/*
 * reordering.c:299-362 (sym) == :145-208 (unsym): from the partition vector
 * to the permuted matrix.  In/out exactly like the reference:
 *   in : I,J,V (any entry order), rowIdx (row-length prefix sums), W, nParts
 *   out: newI,newJ,newV (row-sorted, per-row entry order kept), rowIdx
 *        (updated in place), numInRow (new lengths), numInRow2 (in-window
 *        count per new row; must come in zeroed like the reader's calloc),
 *        partBoundary[0..nParts], reorderList[old] = new.
 */
void orc_reorder(int n, int totalNum, int nParts, int W, const uint32_t *partVec,
                 const int *I, const int *J, const double *V, int *rowIdx,
                 int *numInRow, int *numInRow2_out, int *partBoundary, int *reorderList,
                 int *newI, int *newJ, double *newV)
{
    int *partSize = (int *)calloc((size_t)nParts, sizeof(int));
    int *partBias = (int *)calloc((size_t)nParts + 1, sizeof(int));
    int *partFilled = (int *)calloc((size_t)nParts, sizeof(int));
    int *cSame = (int *)calloc((size_t)n + 1, sizeof(int)); /* the local numInRow2, :254 */
    for (int i = 0; i < n; ++i) partSize[partVec[i]] += 1;               /* :301-303 */
    partBias[0] = 0;
    for (int i = 1; i < nParts + 1; ++i) partBias[i] = partBias[i - 1] + partSize[i - 1];
    for (int i = 0; i < n; ++i) {                                        /* :312-317 */
        reorderList[i] = partFilled[partVec[i]] + partBias[partVec[i]];
        partFilled[partVec[i]] += 1;
        numInRow[i] = 0;
    }
    for (int i = 0; i <= nParts; ++i) partBoundary[i] = partBias[i];     /* :319-321 */
    for (int i = 0; i < totalNum; ++i)                                   /* :327-331 */
        if (partVec[J[i]] == partVec[I[i]]) cSame[I[i]] += 1;
    /* sortRordrList, reordering.c:18-39 */
    orc_rowS *vec = (orc_rowS *)malloc((size_t)n * sizeof(orc_rowS));
    for (int i = 0; i < n; ++i) {
        vec[reorderList[i]].idx = (unsigned)i;
        vec[reorderList[i]].nonzeros = (unsigned)cSame[i];
    }
    for (int p = 0; p < nParts; ++p)
        qsort(&vec[partBoundary[p]], (size_t)(partBoundary[p + 1] - partBoundary[p]),
              sizeof(orc_rowS), orc_rowS_cmp);
    for (int i = 0; i < n; ++i) reorderList[vec[i].idx] = i;
    free(vec);
    /* new row lengths and pointers, :335-345 */
    int *len = (int *)calloc((size_t)n, sizeof(int));
    for (int i = 0; i < n; ++i) len[reorderList[i]] += rowIdx[i + 1] - rowIdx[i];
    rowIdx[0] = 0;
    for (int i = 1; i <= n; ++i) rowIdx[i] = rowIdx[i - 1] + len[i - 1];
    free(len);
    memset(numInRow, 0, (size_t)n * sizeof(int));
    /* scatter, :348-362 */
    for (int i = 0; i < totalNum; ++i) {
        int tI = reorderList[I[i]];
        int tJ = reorderList[J[i]];
        int idx = rowIdx[tI] + numInRow[tI];
        newI[idx] = tI; newJ[idx] = tJ; newV[idx] = V[i];
        numInRow[tI] += 1;
        int partStart = partBoundary[partVec[I[i]]];
        int partEnd = partStart + W;
        if (tJ >= partStart && tJ < partEnd) numInRow2_out[tI] += 1;
    }
    free(partSize); free(partBias); free(partFilled); free(cSame);
}

/* reordering.c:380-391 */
void orc_vector_reorder(int n, const double *v, double *vr, const int *list)
{
    for (int i = 0; i < n; ++i) vr[list[i]] = v[i];
}
void orc_vector_recover(int n, const double *vr, double *v, const int *list)
{
    for (int i = 0; i < n; ++i) v[i] = vr[list[i]];
}

/* ------------------------------------------------------------------ */
/* COO -> EHYB (reference layout)                          */
/* ------------------------------------------------------------------ */

/*
 * Pass 1 of convert.c (vecsGenBlockELL :61-146, prefix sums :336-340,
 * sortRordrListFull :8-31, vecsGenER :148-168, ER prefix :348-354).
 * Caller provides arrays sized: width/bias[nParts*W/32], numInRowER[n],
 * reorderListER[n], rowVecER[n], widthER/biasER[n/32+1], longRow[n].
 * Deviations from the reference, all on inputs where it aborts or is
 * undefined (SURVEY B-3, B-14, B-19): a used slice of width 0 is legal,
 * numOfRowER == 0 is legal, the long-row scan stops at the partition end
 * and long rows are reported in longRow[0..nLong).
 * out[0]=sizeBlockELL out[1]=sizeER out[2]=numOfRowER out[3]=toER out[4]=nLong
 */
void orc_convert_plan(int n, int nParts, int W, const int *partBoundary,
                      const int *numInRow, const int *numInRow2,
                      int16_t *widthELL, int *biasELL,
                      int *numInRowER, int *reorderListER, int *rowVecER,
                      int16_t *widthER, int *biasER, int *realStart, int *longRow,
                      long long *out)
{
    int blockPerPart = W / ORC_WARP;
    int numOfRowER = 0, nLong = 0;
    long long toER = 0;
    memset(numInRowER, 0, (size_t)n * sizeof(int));
    for (int p = 0; p < nParts; ++p) {
        int partStart = partBoundary[p], partEnd = partBoundary[p + 1];
        int rs = partStart;
        while (rs < partEnd && numInRow2[rs] > ORC_LONG) {      /* :92-101 */
            longRow[nLong++] = rs;
            rs += 1;
        }
        realStart[p] = rs;
        for (int it = 0; it < blockPerPart; ++it) {             /* :107-127 */
            int blockStart = partStart + it * ORC_WARP;
            int16_t numCols = 0;
            for (int row = blockStart; row < blockStart + ORC_WARP && row < partEnd; ++row) {
                if (row >= rs) {
                    if (numInRow2[row] > numCols) numCols = (int16_t)numInRow2[row];
                    if (numInRow2[row] != numInRow[row]) {
                        numOfRowER += 1;
                        numInRowER[row] = numInRow[row] - numInRow2[row];
                        toER += numInRowER[row];
                    }
                }
            }
            widthELL[it + blockPerPart * p] = numCols;
        }
        for (int row = partStart + blockPerPart * ORC_WARP; row < partEnd; ++row) { /* :128-134 */
            if (row < rs) continue; /* a long row beyond the window stays a long row */
            numOfRowER += 1;
            numInRowER[row] += numInRow[row];
            toER += numInRowER[row];
        }
    }
    long long sizeELL = 0;
    for (int i = 0; i < blockPerPart * nParts; ++i) {            /* :336-340 */
        biasELL[i] = (int)sizeELL;
        sizeELL += ORC_WARP * (long long)widthELL[i];
    }
    /* sortRordrListFull :8-31 */
    orc_rowS *vec = (orc_rowS *)malloc((size_t)n * sizeof(orc_rowS));
    for (int i = 0; i < n; ++i) { vec[i].idx = (unsigned)i; vec[i].nonzeros = (unsigned)numInRowER[i]; }
    qsort(vec, (size_t)n, sizeof(orc_rowS), orc_rowS_cmp);
    for (int i = 0; i < n; ++i) reorderListER[vec[i].idx] = i;
    free(vec);
    int blockNumER = (numOfRowER + ORC_WARP - 1) / ORC_WARP;     /* :343 */
    for (int i = 0; i < blockNumER; ++i) { widthER[i] = 0; biasER[i] = 0; }
    for (int i = 0; i < n; ++i) {                                /* :154-166 */
        if (numInRowER[i] > 0) {
            int loc = reorderListER[i];
            rowVecER[loc] = i;
            int w = loc / ORC_WARP;
            if (numInRowER[i] > widthER[w]) widthER[w] = (int16_t)numInRowER[i];
        }
    }
    long long sizeER = 0;
    for (int i = 0; i < blockNumER; ++i) {                       /* :348-354 */
        biasER[i] = (int)sizeER;
        sizeER += ORC_WARP * (long long)widthER[i];
    }
    out[0] = sizeELL; out[1] = sizeER; out[2] = numOfRowER; out[3] = toER; out[4] = nLong;
}

/*
 * Pass 2: COO2EHYBCore, convert.c:170-311.  val/col arrays must come in
 * zeroed (the reference callocs them, :341-342, :356-357).  Returns
 * wasteElement (:310), or -1 on one of the reference's consistency aborts.
 */
long long orc_convert_fill(int n, int nParts, int W, const int *partBoundary,
                           const int *rowIdx, const int *numInRow,
                           const int *I, const int *J, const double *V,
                           const int16_t *widthELL, const int *biasELL,
                           const int *numInRowER, const int *reorderListER, int numOfRowER,
                           const int *rowVecER, const int16_t *widthER, const int *biasER,
                           const int *realStart,
                           int16_t *colELL, double *valELL, int *colER, double *valER)
{
    int blockPerPart = W / ORC_WARP;
    long long waste = 0;
    (void)n;
    for (int blockIdx = 0; blockIdx < nParts * blockPerPart; ++blockIdx) {
        int wE = widthELL[blockIdx], bE = biasELL[blockIdx];
        int p = blockIdx / blockPerPart;
        int partStart = partBoundary[p], partEnd = partBoundary[p + 1];
        int rs = realStart[p];
        int fetchEnd = partStart + W;
        int blockStart = partStart + ORC_WARP * (blockIdx % blockPerPart);
        for (int i = 0; i < ORC_WARP; ++i) {
            int wrE = 0, wrR = 0;
            int row = blockStart + i;
            if (row >= rs && row < partEnd) {
                int bR = 0, laneR = 0, wR;
                if (reorderListER[row] < numOfRowER) {           /* :224-233 */
                    int loc = reorderListER[row];
                    if (rowVecER[loc] != row) return -1;
                    bR = biasER[loc / ORC_WARP];
                    laneR = loc % ORC_WARP;
                    wR = widthER[loc / ORC_WARP];
                } else {
                    if (numInRowER[row] > 0) return -1;
                    wR = -1;
                }
                for (int j = 0; j < numInRow[row]; ++j) {        /* :241-268 */
                    int t = j + rowIdx[row];
                    if (I[t] != row) return -1;
                    if (J[t] < fetchEnd && J[t] >= partStart) {
                        colELL[bE + i + wrE * ORC_WARP] = (int16_t)(J[t] - partStart);
                        valELL[bE + i + wrE * ORC_WARP] = V[t];
                        wrE += 1;
                        if (wrE > wE) return -1;
                    } else {
                        if (wR < 0 || wrR >= wR) return -1;
                        colER[bR + laneR + wrR * ORC_WARP] = J[t];
                        valER[bR + laneR + wrR * ORC_WARP] = V[t];
                        wrR += 1;
                    }
                }
                waste += wE - wrE;                               /* :269-274 zero pad */
            } else {
                waste += wE;                                     /* :277-281 */
            }
        }
        if (blockIdx % blockPerPart == 0) {                      /* :285-306 rows beyond window */
            for (int row = partStart + ORC_WARP * blockPerPart; row < partEnd; ++row) {
                if (row < rs) continue;
                int loc = reorderListER[row];
                if (rowVecER[loc] != row) return -1;
                int bR = biasER[loc / ORC_WARP], laneR = loc % ORC_WARP;
                for (int j = 0; j < numInRow[row]; ++j) {
                    colER[bR + laneR + j * ORC_WARP] = J[rowIdx[row] + j];
                    valER[bR + laneR + j * ORC_WARP] = V[rowIdx[row] + j];
                }
            }
        }
    }
    return waste;
}

/* ------------------------------------------------------------------ */
/* kernel semantics                                       */
/* ------------------------------------------------------------------ */

/*
 * CPU emulation of kernelCachedBlockedELL + vecReorderER (+ the intended
 * long-row kernel): kernel.cu:137-163 (ELL phase, window zeroing :139-140
 * generalised to every row beyond the window, B-15), :176-189 (ER phase,
 * first-launch semantics, B-1), :69-77 (scatter-add), :43-67 (long rows,
 * intended behaviour, B-3: whole row, columns from J).
 * use_fma != 0 accumulates with fma() in ascending k, which is what nvcc
 * emits for `dot += val*x` (DFMA): the CUDA product is compared bit-for-bit
 * against that variant.
 */
void orc_emulate(int n, int nParts, int W, const int *partBoundary,
                 const int16_t *widthELL, const int *biasELL,
                 const int16_t *colELL, const double *valELL,
                 int numOfRowER, const int *rowVecER, const int16_t *widthER, const int *biasER,
                 const int *colER, const double *valER,
                 int nLong, const int *longRow, const int *rowIdx, const int *J, const double *V,
                 const double *x, double *y, int use_fma)
{
    int S = W / ORC_WARP;
    (void)n;
    for (int p = 0; p < nParts; ++p) {
        int ps = partBoundary[p], pe = partBoundary[p + 1];
        for (int s = 0; s < S; ++s) {
            int w = widthELL[p * S + s], b = biasELL[p * S + s];
            for (int l = 0; l < ORC_WARP; ++l) {
                int r = ps + ORC_WARP * s + l;
                if (r >= pe) continue;
                double dot = 0;
                for (int k = 0; k < w; ++k) {
                    double v = valELL[b + ORC_WARP * k + l];
                    double xv = x[ps + colELL[b + ORC_WARP * k + l]];
                    dot = use_fma ? fma(v, xv, dot) : dot + v * xv;
                }
                y[r] = dot;
            }
        }
        for (int r = ps + W; r < pe; ++r) y[r] = 0;
    }
    for (int q = 0; q < numOfRowER; ++q) {
        int w = widthER[q / ORC_WARP], b = biasER[q / ORC_WARP], l = q % ORC_WARP;
        double dot = 0;
        for (int k = 0; k < w; ++k) {
            double v = valER[b + l + ORC_WARP * k];
            double xv = x[colER[b + l + ORC_WARP * k]];
            dot = use_fma ? fma(v, xv, dot) : dot + v * xv;
        }
        y[rowVecER[q]] += dot;
    }
    for (int i = 0; i < nLong; ++i) {
        int r = longRow[i];
        double dot = 0;
        for (int k = rowIdx[r]; k < rowIdx[r + 1]; ++k)
            dot = use_fma ? fma(V[k], x[J[k]], dot) : dot + V[k] * x[J[k]];
        y[r] += dot;
    }
}

/* ------------------------------------------------------------------ */
/* CPU baseline: CSR SpMV over the reference's own arrays              */
/* ------------------------------------------------------------------ */

/* plain CSR over rowIdx/J/V (spmv.h:24-29), OpenMP
 * static rows, fp64.  This is the reported `cpu_baseline`, and y_ref of the
 * accuracy gate when run on a row-sorted COO. */
void orc_csr_spmv(int n, const int *rowIdx, const int *J, const double *V,
                  const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r) {
        double s = 0;
        for (int k = rowIdx[r]; k < rowIdx[r + 1]; ++k) s += V[k] * x[J[k]];
        y[r] = s;
    }
}

/* 64-bit row pointers (local blocks of config 5 stay < 2^31 nnz, but the
 * multi-rank emulation concatenates them). */
void orc_csr_spmv64(long long n, const long long *rowPtr, const int *J, const double *V,
                    const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (long long r = 0; r < n; ++r) {
        double s = 0;
        for (long long k = rowPtr[r]; k < rowPtr[r + 1]; ++k) s += V[k] * x[J[k]];
        y[r] = s;
    }
}

/* (|A||x|)_row for the accuracy gate of BASELINE.json:
 * |y - y_ref| <= 1e-12 * (|A||x|) per row. */
void orc_csr_abs_spmv(int n, const int *rowIdx, const int *J, const double *V,
                      const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r) {
        double s = 0;
        for (int k = rowIdx[r]; k < rowIdx[r + 1]; ++k) s += fabs(V[k]) * fabs(x[J[k]]);
        y[r] = s;
    }
}

/* Timed loop for the cpu_baseline leg: returns seconds for `iters` products. */
double orc_csr_spmv_timed(int n, const int *rowIdx, const int *J, const double *V,
                          const double *x, double *y, int warmup, int iters)
{
    for (int i = 0; i < warmup; ++i) orc_csr_spmv(n, rowIdx, J, V, x, y);
#ifdef _OPENMP
    double t0 = omp_get_wtime();
    for (int i = 0; i < iters; ++i) orc_csr_spmv(n, rowIdx, J, V, x, y);
    return omp_get_wtime() - t0;
#else
    return -1.0;
#endif
}

/* The driver's self-check: solver_test.c:7-29.  out[0]=sum|d|, out[1]=sum
 * |d|/ampl, returns the number of rows over threshold. */
int orc_compare(const double *yResult, const double *y, double threshold, int n, double *out)
{
    double avgdiff = 0, avgampl = 0;
    int k = 0;
    for (int i = 0; i < n; ++i) {
        double d = fabs(y[i] - yResult[i]);
        double ampl = fmin(fabs(y[i]), fabs(yResult[i]));
        if (d > ampl * threshold) k++;
        avgdiff += d;
        if (ampl > 0) avgampl += d / ampl;
    }
    out[0] = avgdiff; out[1] = avgampl;
    return k;
}

/* ------------------------------------------------------------------------- */
/* BASELINE.json config 5 (27-point stencil 512^3): the matrix is too large to   */
/* hold as a CSR next to the product under test, and the reference cannot run it */
/* at all, so the check vector is evaluated in closed form  */
/* from the generator's definition (diag 26, off -1, 27-point      */
/* neighbourhood clipped at the boundary; same matrix as gen_stencil27_lower +   */
/* solver_test.c:127-265's symmetric expansion, checked against each other in    */
/* tests/test_oracle_pinned.py) for x a fixed hash of the natural row index.     */
/* ------------------------------------------------------------------------- */

/* x(i): splitmix-style hash of the natural index, in (-0.1, 0.1); the same
 * function as multigpu.x_of_global (numpy), bit for bit */
static inline double orc_x_of_global(unsigned long long i)
{
    unsigned long long z = i * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL;
    z ^= z >> 29;
    z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 32;
    return ((double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 0.2;
}

void orc_x_of_global_fill(long long count, const long long *ids, double scale, double shift, double *x)
{
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < count; ++k) x[k] = orc_x_of_global((unsigned long long)ids[k]) * scale + shift;
}

/* y_ref and |A||x| of the rows `ids` (natural indices (z*ny + y)*nx + x) for
 * x_j = orc_x_of_global(j) * scale + shift; entries summed in ascending column
 * order like the CSR baseline. */
void orc_stencil27_rows_product(int nx, int ny, int nz, long long count, const long long *ids, double scale, double shift,
                                double *yref, double *absAx)
{
    const long long plane = (long long)nx * ny;
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < count; ++k) {
        const long long g = ids[k];
        const int x = (int)(g % nx), y = (int)((g / nx) % ny), z = (int)(g / plane);
        double s = 0.0, a = 0.0;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (x + dx < 0 || x + dx >= nx || y + dy < 0 || y + dy >= ny || z + dz < 0 || z + dz >= nz) continue;
                    const long long c = g + dz * plane + (long long)dy * nx + dx;
                    const double v = c == g ? 26.0 : -1.0;
                    const double xv = orc_x_of_global((unsigned long long)c) * scale + shift;
                    s += v * xv;
                    a += fabs(v) * fabs(xv);
                }
        yref[k] = s;
        absAx[k] = a;
    }
}
