// lzma_dec.cu -- host-side LZMA1 raw-stream decoder for the legacy MLV_VIDEO_CLASS_FLAG_LZMA payloads.
//
// The reference hands such frames to the LZMA SDK on the CPU (main.c:598-616: uint32 unpacked size, 5 property
// bytes, then the raw LZMA stream; LzmaUncompress with LZMA_FINISH_ANY) and then unpacks the result with
// dng_get_image_data.  The codec is a serial range coder, so it stays on the host here too (SURVEY.md 8(f) rank 3:
// "LZMA stays CPU"); only the unpack and everything after it run on the GPU.  This is a from-scratch decoder
// written against the published LZMA format (state machine of 12 states, literal / match / rep0-3 / short-rep
// packets, 11-bit adaptive probabilities, 32-bit range coder), decoding straight into the flat output buffer --
// the "dictionary" is the output itself, since a frame is decoded in one piece.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/mlvfs_b200.h"   // status codes only: this file is plain host C++ (tests also build it with g++)

namespace {

constexpr int kNumBitModelTotalBits = 11;
constexpr uint16_t kProbInit = (1u << kNumBitModelTotalBits) / 2;
constexpr int kNumMoveBits = 5;
constexpr uint32_t kTopValue = 1u << 24;

struct RangeDecoder {
    const uint8_t *p, *end;
    uint32_t range = 0xFFFFFFFFu, code = 0;
    bool overrun = false;

    uint8_t next()
    {
        if (p < end) return *p++;
        overrun = true;
        return 0;
    }
    bool init()
    {
        if (next() != 0) return false;                    // the first byte of a range-coded stream is always 0
        for (int i = 0; i < 4; i++) code = (code << 8) | next();
        return !overrun && code != range;
    }
    void normalize()
    {
        if (range < kTopValue) { range <<= 8; code = (code << 8) | next(); }
    }
    unsigned bit(uint16_t *prob)
    {
        const uint32_t bound = (range >> kNumBitModelTotalBits) * *prob;
        unsigned b;
        if (code < bound) { range = bound; *prob += ((1u << kNumBitModelTotalBits) - *prob) >> kNumMoveBits; b = 0; }
        else { range -= bound; code -= bound; *prob -= *prob >> kNumMoveBits; b = 1; }
        normalize();
        return b;
    }
    uint32_t direct(int nbits)
    {
        uint32_t r = 0;
        while (nbits-- > 0) {
            range >>= 1;
            code -= range;
            const uint32_t t = 0u - (code >> 31);         // all ones when the subtraction went negative
            code += range & t;
            r = (r << 1) + t + 1;
            normalize();
        }
        return r;
    }
};

unsigned tree(RangeDecoder &rc, uint16_t *probs, int nbits)
{
    unsigned m = 1;
    for (int i = 0; i < nbits; i++) m = (m << 1) + rc.bit(&probs[m]);
    return m - (1u << nbits);
}

unsigned tree_reverse(RangeDecoder &rc, uint16_t *probs, int nbits)
{
    unsigned m = 1, sym = 0;
    for (int i = 0; i < nbits; i++) {
        const unsigned b = rc.bit(&probs[m]);
        m = (m << 1) + b;
        sym |= b << i;
    }
    return sym;
}

struct LenDecoder {
    uint16_t choice = kProbInit, choice2 = kProbInit;
    uint16_t low[16][8], mid[16][8], high[256];
    LenDecoder()
    {
        for (auto &r : low) for (auto &v : r) v = kProbInit;
        for (auto &r : mid) for (auto &v : r) v = kProbInit;
        for (auto &v : high) v = kProbInit;
    }
    unsigned decode(RangeDecoder &rc, unsigned pos_state)
    {
        if (rc.bit(&choice) == 0) return tree(rc, low[pos_state], 3);
        if (rc.bit(&choice2) == 0) return 8 + tree(rc, mid[pos_state], 3);
        return 16 + tree(rc, high, 8);
    }
};

constexpr int kNumStates = 12, kNumLenToPosStates = 4, kEndPosModelIndex = 14, kNumFullDistances = 1 << (kEndPosModelIndex >> 1);
constexpr int kNumAlignBits = 4, kMatchMinLen = 2;

}  // namespace

// Decodes up to *dst_len bytes; on return *dst_len = bytes produced.  Returns 0 when the stream decoded cleanly up
// to the requested size or to its end marker, < 0 for a corrupt / truncated stream or bad properties.
int mlvb_lzma_decode(uint8_t *dst, size_t *dst_len, const uint8_t *src, size_t src_len, const uint8_t props[5])
{
    const size_t out_size = *dst_len;
    *dst_len = 0;
    unsigned d = props[0];
    if (d >= 9 * 5 * 5) return MLVB_ERR_ARG;
    const int lc = d % 9; d /= 9;
    const int lp = d % 5, pb = d / 5;
    // props[1..4] is the dictionary size; distances are checked against what has been produced instead

    RangeDecoder rc{src, src + src_len};
    if (!rc.init()) return MLVB_ERR_ARG;

    std::vector<uint16_t> literal((size_t)0x300 << (lc + lp), kProbInit);
    uint16_t is_match[kNumStates][16], is_rep[kNumStates], is_rep_g0[kNumStates], is_rep_g1[kNumStates], is_rep_g2[kNumStates],
        is_rep0_long[kNumStates][16];
    uint16_t pos_slot[kNumLenToPosStates][64], pos_special[1 + kNumFullDistances - kEndPosModelIndex], pos_align[1 << kNumAlignBits];
    for (auto &r : is_match) for (auto &v : r) v = kProbInit;
    for (auto &r : is_rep0_long) for (auto &v : r) v = kProbInit;
    for (int i = 0; i < kNumStates; i++) is_rep[i] = is_rep_g0[i] = is_rep_g1[i] = is_rep_g2[i] = kProbInit;
    for (auto &r : pos_slot) for (auto &v : r) v = kProbInit;
    for (auto &v : pos_special) v = kProbInit;
    for (auto &v : pos_align) v = kProbInit;
    LenDecoder len_dec, rep_len_dec;

    uint32_t rep0 = 0, rep1 = 0, rep2 = 0, rep3 = 0;
    unsigned state = 0;
    size_t pos = 0;
    while (pos < out_size) {
        if (rc.overrun) return MLVB_ERR_ARG;
        const unsigned pos_state = (unsigned)pos & ((1u << pb) - 1);
        if (rc.bit(&is_match[state][pos_state]) == 0) {
            // literal
            const unsigned prev = pos ? dst[pos - 1] : 0;
            const unsigned lit_state = (((unsigned)pos & ((1u << lp) - 1)) << lc) + (prev >> (8 - lc));
            uint16_t *probs = &literal[(size_t)0x300 * lit_state];
            unsigned sym = 1;
            if (state >= 7) {
                unsigned match_byte = dst[pos - rep0 - 1];
                do {
                    const unsigned match_bit = (match_byte >> 7) & 1;
                    match_byte <<= 1;
                    const unsigned b = rc.bit(&probs[((1 + match_bit) << 8) + sym]);
                    sym = (sym << 1) | b;
                    if (match_bit != b) break;
                } while (sym < 0x100);
            }
            while (sym < 0x100) sym = (sym << 1) | rc.bit(&probs[sym]);
            dst[pos++] = (uint8_t)sym;
            state = state < 4 ? 0 : (state < 10 ? state - 3 : state - 6);
            continue;
        }
        unsigned len;
        if (rc.bit(&is_rep[state]) != 0) {
            if (pos == 0) return MLVB_ERR_ARG;
            if (rc.bit(&is_rep_g0[state]) == 0) {
                if (rc.bit(&is_rep0_long[state][pos_state]) == 0) {          // short rep: one byte from rep0
                    state = state < 7 ? 9 : 11;
                    dst[pos] = dst[pos - rep0 - 1];
                    pos++;
                    continue;
                }
            } else {
                uint32_t dist;
                if (rc.bit(&is_rep_g1[state]) == 0) dist = rep1;
                else {
                    if (rc.bit(&is_rep_g2[state]) == 0) dist = rep2;
                    else { dist = rep3; rep3 = rep2; }
                    rep2 = rep1;
                }
                rep1 = rep0;
                rep0 = dist;
            }
            len = rep_len_dec.decode(rc, pos_state);
            state = state < 7 ? 8 : 11;
        } else {
            rep3 = rep2; rep2 = rep1; rep1 = rep0;
            len = len_dec.decode(rc, pos_state);
            state = state < 7 ? 7 : 10;
            const unsigned len_state = len < kNumLenToPosStates - 1 ? len : kNumLenToPosStates - 1;
            const unsigned slot = tree(rc, pos_slot[len_state], 6);
            if (slot < 4) rep0 = slot;
            else {
                const int nbits = (int)(slot >> 1) - 1;
                uint32_t dist = (2 | (slot & 1)) << nbits;
                if (slot < (unsigned)kEndPosModelIndex) dist += tree_reverse(rc, pos_special + dist - slot, nbits);
                else {
                    dist += rc.direct(nbits - kNumAlignBits) << kNumAlignBits;
                    dist += tree_reverse(rc, pos_align, kNumAlignBits);
                }
                rep0 = dist;
            }
            if (rep0 == 0xFFFFFFFFu) break;                                  // end-of-stream marker
        }
        if ((size_t)rep0 >= pos) return MLVB_ERR_ARG;                         // distance reaches before the start
        len += kMatchMinLen;
        size_t n = len;
        if (n > out_size - pos) n = out_size - pos;                           // LZMA_FINISH_ANY: stop when the buffer is full
        const uint8_t *from = dst + pos - rep0 - 1;
        for (size_t i = 0; i < n; i++) dst[pos + i] = from[i];                // may overlap forwards (run replication)
        pos += n;
    }
    if (rc.overrun) return MLVB_ERR_ARG;
    *dst_len = pos;
    return MLVB_OK;
}
