/* oracle/orc_unpack.c -- TEST INFRASTRUCTURE. Bit-unpack, restating dng.c:813-872. */
#include "oracle.h"

/*
 * Pixel i occupies bits [i*bpp, (i+1)*bpp) of the stream obtained by reading the payload as
 * 16-bit little-endian words and concatenating them most-significant-bit first (raw.h:41-79).
 * The reference does a 32-bit load of words k,k+1 and a rotate (dng.c:836-840); the arithmetic
 * below is the same thing written as a shift of the 32-bit big-endian pair.
 * offset/max_size select a byte sub-range of the unpacked frame; `packed` then points at the
 * word that holds the first requested pixel (main.c:692: the caller seeks to pixel_start_address).
 */
size_t orc_unpack(const uint16_t *packed, uint8_t *out, long offset, size_t max_size, int bpp)
{
    uint32_t first = (uint32_t)(offset > 0 ? offset : 0) / 2;       /* dng.c:815 */
    uint32_t first_word = first * (uint32_t)bpp / 16;               /* dng.c:816 */
    size_t skip = offset < 0 ? (size_t)(-offset) : 0;
    size_t out_bytes = max_size - skip;                             /* dng.c:817 */
    uint32_t mask = (1u << bpp) - 1;
    uint16_t *dst = (uint16_t *)(out + skip + offset % 2);          /* dng.c:823 */
    uint32_t n = (uint32_t)(out_bytes / 2);
    for (uint32_t j = 0; j < n; j++) {
        uint32_t bit = (first + j) * (uint32_t)bpp;
        uint32_t k = bit / 16 - first_word, s = bit % 16;
        /* the reference's 32-bit load also touches word k+1 when the pixel ends inside word k
           (hence packed_size = (px+2)*bpp/16 at main.c:579); those bits are masked away, so the
           restatement simply does not read them */
        uint32_t pair = ((uint32_t)packed[k] << 16) | (s + bpp > 16 ? packed[k + 1] : 0u);
        dst[j] = (uint16_t)((pair >> (32 - bpp - s)) & mask);
    }
    return max_size;
}
