/* oracle/orc_cs.c -- TEST INFRASTRUCTURE. Median chroma smoothing, restating cs.c:37-84 and the
 * chroma_smooth.c:22-71 template (instantiated for uint16 in cs.c and uint32 in hdr.c:1488-1500). */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* All EV arithmetic wraps mod 2^32 like the compiled reference does (raw2ev[black] is INT_MIN). */
static inline int wadd(int a, int b) { return (int)((uint32_t)a + (uint32_t)b); }
static inline int wsub(int a, int b) { return (int)((uint32_t)a - (uint32_t)b); }

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* The reference uses fixed compare-exchange networks (opt_med.h:41-46,107-116,129-168); they return
   the exact middle order statistic of 5 / 9 / 25 values, so a sort gives the same value. */
static int median_odd(int *v, int n)
{
    qsort(v, (size_t)n, sizeof(int), cmp_int);
    return v[n / 2];
}

static inline int clamp_ev(int e)
{
    return e < 0 ? 0 : (e > 14 * ORC_EV_RES - 1 ? 14 * ORC_EV_RES - 1 : e);
}

#define CS_BODY(T)                                                                               \
    int reach = (method == 5) ? 4 : 2;                                                           \
    for (int y = 4; y < h - 5; y += 2) {                     /* chroma_smooth.c:26 */            \
        for (int x = 4; x < w - 4; x += 2) {                 /* chroma_smooth.c:28 */            \
            int g1 = (int)inp[x + 1 + y * w], g2 = (int)inp[x + (y + 1) * w];                    \
            int ge = wadd(raw2ev[g1], raw2ev[g2]) / 2;                                           \
            if (ge < 2 * ORC_EV_RES) continue;               /* :35 */                           \
            int mr[25], mb[25], k = 0;                                                           \
            for (int i = -reach; i <= reach; i += 2)                                             \
                for (int j = -reach; j <= reach; j += 2) {                                       \
                    if (method == 2 && abs(i) + abs(j) == 4) continue;   /* :45-48 */            \
                    int r  = (int)inp[x + i     + (y + j) * w];                                  \
                    int a1 = (int)inp[x + i + 1 + (y + j) * w];                                  \
                    int a2 = (int)inp[x + i     + (y + j + 1) * w];                              \
                    int b  = (int)inp[x + i + 1 + (y + j + 1) * w];                              \
                    int gq = wadd(raw2ev[a1], raw2ev[a2]) / 2;                                   \
                    mr[k] = wsub(raw2ev[r], gq);                                                 \
                    mb[k] = wsub(raw2ev[b], gq);                                                 \
                    k++;                                                                         \
                }                                                                                \
            int dr = median_odd(mr, k), db = median_odd(mb, k);                                  \
            if (wadd(ge, dr) <= ORC_EV_RES) continue;        /* :63 */                           \
            if (wadd(ge, db) <= ORC_EV_RES) continue;        /* :64 */                           \
            out[x + y * w]           = (T)(ev2raw[clamp_ev(wadd(ge, dr))] + black);              \
            out[x + 1 + (y + 1) * w] = (T)(ev2raw[clamp_ev(wadd(ge, db))] + black);              \
        }                                                                                        \
    }

static void cs_u16(const uint16_t *inp, uint16_t *out, int w, int h, int method,
                   const int *raw2ev, const int *ev2raw, int black)
{
    CS_BODY(uint16_t)
}

void orc_chroma_smooth_u32(const uint32_t *inp, uint32_t *out, int w, int h, int method,
                           const int *raw2ev, const int *ev2raw)
{
    const int black = 0;
    CS_BODY(uint32_t)
}

/* cs.c:49-84: smooth in place from an untouched copy; unknown method leaves the frame alone */
void orc_chroma_smooth_u16(uint16_t *img, int w, int h, int black, int method)
{
    const int *raw2ev = orc_raw2ev(black);
    if (!raw2ev || (method != 2 && method != 3 && method != 5)) return;
    size_t bytes = (size_t)w * h * sizeof(uint16_t);
    uint16_t *copy = malloc(bytes);
    if (!copy) return;
    memcpy(copy, img, bytes);
    cs_u16(copy, img, w, h, method, raw2ev, orc_ev2raw(), black);
    free(copy);
}
