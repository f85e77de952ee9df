/* ovfstream.h -- device format of a large overflow list (internal; see ovfstream.c). */
#ifndef EHYB_OVFSTREAM_H
#define EHYB_OVFSTREAM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ehyb_ovfstream {
    int64_t count, nGroups, nSeg, hubRefs, deviceBytes;
    int nHub;
    uint32_t *col;      /* [count] column, or 0x80000000 | hub index */
    uint32_t *grp;      /* [nGroups][2] {segment of the group's first entry, new-row mask} */
    int32_t *rowOfSeg;  /* [nSeg] */
    int32_t *hubCols;   /* [nHub] */
} ehyb_ovfstream;

/* row[] sorted ascending; values stay in the layout's ovfVal (same order). */
int ehyb_ovfstream_build(int64_t count, const int32_t *row, const int32_t *col, int64_t ncols, int hubCap, ehyb_ovfstream *out);
void ehyb_ovfstream_free(ehyb_ovfstream *s);

#ifdef __cplusplus
}
#endif
#endif
