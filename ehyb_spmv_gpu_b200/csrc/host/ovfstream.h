/* ovfstream.h -- device format of a large overflow list (internal; see ovfstream.c). */
#ifndef EHYB_OVFSTREAM_H
#define EHYB_OVFSTREAM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* one tile = tileGroups x 32 consecutive entries of the row-sorted list, packed into ONE record a
 * single bulk copy moves into shared memory (E = 32 * tileGroups):
 *   [0, 8E)                 double   val[E]
 *   [8E, 12E)               uint32   col[E]        column, or 0x80000000 | hub index
 *   [12E, 12E + 8 TG)       {uint32 seg0, uint32 mask}[TG]   per 32 entries: segment of the first
 *                                    entry, bit j = entry j starts a new row
 *   [12E + 8 TG, +16)       uint32   flags (bit 0: the first entry continues the previous tile's
 *                                    row, bit 1: the next tile continues this tile's last row), 3 x 0 */
#define EHYB_OVF_TILE_BYTES(tg) (12 * 32 * (tg) + 8 * (tg) + 16)
#define EHYB_OVF_HUB_BIT 0x80000000u

typedef struct ehyb_ovfstream {
    int64_t count, nTiles, nSeg, hubRefs, deviceBytes;
    int nHub, tileGroups, tileBytes;
    unsigned char *tiles; /* [nTiles * tileBytes] */
    int32_t *rowOfSeg;    /* [nSeg]: the row of segment s; -1 for the segment of the padding entries */
    int32_t *hubCols;     /* [nHub] */
    int32_t *carryRow;    /* [2 * nTiles]: row of a tile's head / tail carry slot, -1 = slot unused */
    /* the carry slots of one row are a RUN of consecutive slots; the fix-up kernel adds the nRunsShort
     * runs of at most EHYB_OVF_SHORT_RUN slots with one thread each and the long ones (a row that spans
     * hundreds of tiles) with one warp each: runs[] = short runs first, {first slot, slots, row} */
    int64_t nRuns, nRunsShort;
    int32_t *runs;        /* [nRuns][3] */
} ehyb_ovfstream;
#define EHYB_OVF_SHORT_RUN 16

/* row[] sorted ascending (per-row order kept); tileGroups = 4 or 8. */
int ehyb_ovfstream_build(int64_t count, const int32_t *row, const int32_t *col, const double *val, int64_t ncols, int hubCap,
                         int tileGroups, ehyb_ovfstream *out);
void ehyb_ovfstream_free(ehyb_ovfstream *s);

/* COLUMN BLOCKS: the list cut into nBlocks lists by column range (block b = columns [b, b+1) *
 * blockCols), each of them row-sorted with the per-row order kept, each built as a stream of its
 * own into out[b] (count 0 and no arrays for an empty block).  The device runs them one after the
 * other and ADDS into y: all the gathers of a launch fall into one slice of x that stays in L2. */
int ehyb_ovfstream_build_blocked(int64_t count, const int32_t *row, const int32_t *col, const double *val, int64_t ncols, int hubCap,
                                 int tileGroups, int nBlocks, int64_t blockCols, ehyb_ovfstream *out);

#ifdef __cplusplus
}
#endif
#endif
