/* oracle/orc_lj92.c -- TEST INFRASTRUCTURE.  Lossless JPEG (ITU-T T.81 Annex H, "LJ92") for MLV
 * frames: a decoder restating what reference lj92.c:225-271,344-593,650-702 computes, the
 * quadrant de-interleave of main.c:656-668, and an independent encoder used only to GENERATE
 * config-5 test input (the reference's own encoder, lj92.c:1104-1144, is used for cross-checks
 * where oracle/_ref exists).
 *
 * Stream handled (same subset as the reference): SOI, [DHT | SOF3 | other segments]*, SOS, entropy
 * data, one component, one Huffman table, predictors 1..7, 0xFF00 byte stuffing.
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

typedef struct {
    const uint8_t *p, *end;
    uint32_t acc;
    int n;
} bitreader;

static int get_bit(bitreader *br)
{
    if (br->n == 0) {
        uint32_t byte = br->p < br->end ? *br->p++ : 0;
        if (byte == 0xFF && br->p < br->end) br->p++;          /* stuffed 0x00 (lj92.c:356-368) */
        br->acc = byte;
        br->n = 8;
    }
    br->n--;
    return (br->acc >> br->n) & 1;
}

static int get_bits(bitreader *br, int count)
{
    int v = 0;
    while (count--) v = (v << 1) | get_bit(br);
    return v;
}

/* returns 0 on success; *w,*h,*bits from SOF3; out receives w*h samples in stream (tiled) order */
int orc_lj92_decode(const uint8_t *data, int len, uint16_t *out, int cap, int *w, int *h, int *bits)
{
    int ix = 0, width = 0, height = 0, depth = 0, pred = -1, have_table = 0, scan = -1;
    int counts[17] = {0};
    uint8_t vals[256];
    if (len < 4 || data[0] != 0xFF || data[1] != 0xD8) return -1;
    ix = 2;
    while (ix + 4 <= len && scan < 0) {
        if (data[ix] != 0xFF) { ix++; continue; }
        int marker = data[ix + 1];
        int seg = (data[ix + 2] << 8) | data[ix + 3];
        const uint8_t *s = data + ix + 4;
        if (marker == 0xC4) {                                   /* DHT (lj92.c:83-271) */
            int total = 0;
            for (int i = 1; i <= 16; i++) { counts[i] = s[i]; total += counts[i]; }
            if (total > 256) return -1;
            memcpy(vals, s + 17, (size_t)total);
            have_table = 1;
        } else if (marker == 0xC3) {                            /* SOF3 (lj92.c:273-280) */
            depth = s[0];
            height = (s[1] << 8) | s[2];
            width = (s[3] << 8) | s[4];
        } else if (marker == 0xDA) {                            /* SOS (lj92.c:512-519) */
            int ncomp = s[0];
            pred = s[1 + 2 * ncomp];
            scan = ix + 2 + seg;
        } else if (marker == 0xD9) {
            return -1;
        }
        ix += 2 + seg;
    }
    if (scan < 0 || !have_table || width <= 0 || height <= 0 || pred < 1 || pred > 7) return -1;
    if ((long)width * height > cap) return -2;
    *w = width; *h = height; *bits = depth;

    /* canonical code tables (T.81 Annex C / F.2.2.3) */
    int mincode[17], maxcode[18], valptr[17], code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
        valptr[l] = k;
        mincode[l] = code;
        code += counts[l];
        k += counts[l];
        maxcode[l] = counts[l] ? code - 1 : -1;
        code <<= 1;
    }
    bitreader br = {data + scan, data + len, 0, 0};
    for (int y = 0; y < height; y++) {
        uint16_t *row = out + (size_t)y * width, *up = row - width;
        for (int x = 0; x < width; x++) {
            int c = get_bit(&br), l = 1;
            while (l <= 16 && (maxcode[l] < 0 || c > maxcode[l])) { c = (c << 1) | get_bit(&br); l++; }
            if (l > 16) return -3;
            int t = vals[valptr[l] + c - mincode[l]];
            int diff = get_bits(&br, t);
            if (t && diff < (1 << (t - 1))) diff += (int)((~0u) << t) + 1;     /* EXTEND (lj92.c:392-396) */
            int px;
            if (x == 0 && y == 0) px = 1 << (depth - 1);
            else if (y == 0) px = row[x - 1];
            else if (x == 0) px = up[0];
            else {
                int a = row[x - 1], b = up[x], cc = up[x - 1];
                switch (pred) {
                case 1: px = a; break;
                case 2: px = b; break;
                case 3: px = cc; break;
                case 4: px = a + b - cc; break;
                case 5: px = a + ((b - cc) >> 1); break;
                case 6: px = b + ((a - cc) >> 1); break;
                default: px = (a + b) >> 1; break;
                }
            }
            row[x] = (uint16_t)(px + diff);
        }
    }
    return 0;
}

/* main.c:656-668: the stream holds even rows first and, within a row, even columns first */
void orc_lj92_untile(const uint16_t *src, uint16_t *dst, int w, int h)
{
    for (int y = 0; y < h; y++) {
        int dy = ((2 * y) % h) + ((2 * y) / h);
        for (int x = 0; x < w; x++) {
            int dx = ((2 * x) % w) + ((2 * x) / w);
            dst[(size_t)dy * w + dx] = src[(size_t)y * w + x];
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Independent encoder (test-input generator): predictor 6, one Huffman table built from the SSSS
 * histogram with the length-limiting procedure of T.81 Annex K.2.
 */
static int ssss_of(int diff)
{
    int a = diff < 0 ? -diff : diff, s = 0;
    while (a) { s++; a >>= 1; }
    return s;
}

typedef struct { uint8_t *buf; size_t len, cap; uint32_t acc; int n; } bitwriter;

static void put_bits(bitwriter *bw, uint32_t v, int count)
{
    while (count > 0) {
        int take = 8 - bw->n < count ? 8 - bw->n : count;
        bw->acc = (bw->acc << take) | ((v >> (count - take)) & ((1u << take) - 1));
        bw->n += take;
        count -= take;
        if (bw->n == 8) {
            bw->buf[bw->len++] = (uint8_t)bw->acc;
            if ((bw->acc & 0xFF) == 0xFF) bw->buf[bw->len++] = 0;
            bw->acc = 0; bw->n = 0;
        }
    }
}

static int predict6(const uint16_t *img, int w, int x, int y, int depth)
{
    const uint16_t *row = img + (size_t)y * w, *up = row - w;
    if (x == 0 && y == 0) return 1 << (depth - 1);
    if (y == 0) return row[x - 1];
    if (x == 0) return up[0];
    return up[x] + ((row[x - 1] - up[x - 1]) >> 1);
}

/* returns encoded length, or -1; out must hold at least w*h*4 + 1024 bytes */
long orc_lj92_encode(const uint16_t *img, int w, int h, int depth, uint8_t *out, size_t cap)
{
    long freq[18] = {0};
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) freq[ssss_of((int)img[(size_t)y * w + x] - predict6(img, w, x, y, depth))]++;
    freq[17] = 1;                                              /* reserved symbol: no all-ones code */
    int codesize[18] = {0}, others[18];
    for (int i = 0; i < 18; i++) others[i] = -1;
    for (;;) {                                                 /* Annex K.2, figure K.1 */
        int v1 = -1, v2 = -1;
        for (int i = 0; i < 18; i++) if (freq[i] > 0 && (v1 < 0 || freq[i] <= freq[v1])) v1 = i;
        for (int i = 0; i < 18; i++) if (i != v1 && freq[i] > 0 && (v2 < 0 || freq[i] <= freq[v2])) v2 = i;
        if (v2 < 0) break;
        freq[v1] += freq[v2];
        freq[v2] = 0;
        for (codesize[v1]++; others[v1] >= 0;) { v1 = others[v1]; codesize[v1]++; }
        others[v1] = v2;
        for (codesize[v2]++; others[v2] >= 0;) { v2 = others[v2]; codesize[v2]++; }
    }
    int nbits[33] = {0};
    for (int i = 0; i < 18; i++) if (codesize[i]) nbits[codesize[i]]++;
    for (int i = 32; i > 16; i--)                               /* figure K.3: limit to 16 bits */
        while (nbits[i] > 0) {
            int j = i - 2;
            while (nbits[j] == 0) j--;
            nbits[i] -= 2; nbits[i - 1]++; nbits[j + 1] += 2; nbits[j]--;
        }
    { int i = 16; while (nbits[i] == 0) i--; nbits[i]--; }     /* drop the reserved symbol */
    /* symbols in order of increasing code length (ties: increasing value) */
    int order[17], n = 0;
    for (int l = 1; l <= 32; l++) for (int s = 0; s < 17; s++) if (codesize[s] == l) order[n++] = s;
    uint32_t code_of[17] = {0};
    int len_of[17] = {0}, code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
        for (int c = 0; c < nbits[l]; c++) { code_of[order[k]] = (uint32_t)code++; len_of[order[k]] = l; k++; }
        code <<= 1;
    }
    if (cap < (size_t)w * h * 4 + 1024) return -1;
    size_t o = 0;
    out[o++] = 0xFF; out[o++] = 0xD8;
    out[o++] = 0xFF; out[o++] = 0xC3; out[o++] = 0; out[o++] = 11; out[o++] = (uint8_t)depth;
    out[o++] = (uint8_t)(h >> 8); out[o++] = (uint8_t)h; out[o++] = (uint8_t)(w >> 8); out[o++] = (uint8_t)w;
    out[o++] = 1; out[o++] = 0; out[o++] = 0x11; out[o++] = 0;
    out[o++] = 0xFF; out[o++] = 0xC4; out[o++] = 0; out[o++] = (uint8_t)(19 + n); out[o++] = 0;
    for (int l = 1; l <= 16; l++) out[o++] = (uint8_t)nbits[l];
    for (int i = 0; i < n; i++) out[o++] = (uint8_t)order[i];
    out[o++] = 0xFF; out[o++] = 0xDA; out[o++] = 0; out[o++] = 8; out[o++] = 1; out[o++] = 0; out[o++] = 0;
    out[o++] = 6; out[o++] = 0; out[o++] = 0;
    bitwriter bw = {out, o, cap, 0, 0};
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int diff = (int)img[(size_t)y * w + x] - predict6(img, w, x, y, depth);
            int t = ssss_of(diff);
            put_bits(&bw, code_of[t], len_of[t]);
            if (t) put_bits(&bw, (uint32_t)(diff < 0 ? diff + (1 << t) - 1 : diff), t);
        }
    if (bw.n) put_bits(&bw, (1u << (8 - bw.n)) - 1, 8 - bw.n);  /* pad with ones */
    o = bw.len;
    out[o++] = 0xFF; out[o++] = 0xD9;
    return (long)o;
}
