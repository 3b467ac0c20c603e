// lj92_core.cuh -- intra-frame parallel lossless-JPEG ("LJ92") decode: the cooperative programs behind
// the kernels of lj92.cu, written once over a small "context" (thread id, barriers, warp shuffles) so
// that the same source runs as CUDA thread blocks and, for tests, as groups of host threads
// (tests/emu/lj92_emu.cpp).
//
// Replaces reference lj92.c:650-702 (lj92_open / lj92_decode: marker walk lj92.c:83-280 + :595-640,
// Huffman table lj92.c:225-271, bit reader with 0xFF stuffing lj92.c:344-406, predictors
// lj92.c:408-593).  The reference walks the scan serially (a scan has no restart markers).  Here a
// frame is decoded by thousands of threads:
//
//   1. unstuff      the entropy-coded segment is compacted to a clean bit stream (the byte after every
//                   0xFF data byte is dropped, exactly the rule of lj92.c:356-368);
//   2. synchronise  the clean stream is cut into 1024-bit subsequences, one thread each.  Every thread
//                   decodes from a guessed start; Huffman streams re-synchronise after a few symbols, so
//                   iterating  start[i+1] = end[i]  reaches the fixed point -- the true parse, because
//                   start[0] is exact -- in 2-3 rounds (dec_body<0>, then across thread blocks
//                   dec_body<1> and the serial-order check of resolve_body, which guarantees convergence
//                   for ANY stream, however adversarial its code table);
//   3. index        symbol counts are prefix-summed, giving each subsequence its first pixel index;
//   4. write        a last decode pass stores the 16-bit differences at their pixel index (dec_body<2>);
//   5. predict      the predictor recurrence (left / above / above-left) is run as a skewed wavefront:
//                   a warp owns 32 rows, lane l trails lane l-1 by one column and receives "above" with
//                   one shuffle per step; a strip hands its last row to the next through tagged words (predict_strip);
//   6. untile       the quadrant de-interleave of main.c:656-668 (lj92.cu).
//
// All sample arithmetic is modulo 2^16 like the reference's uint16 row buffers; differences are kept
// as int16 (exact for every stream whose samples fit 16 bits, i.e. every valid stream).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LJ_HD __host__ __device__ __forceinline__
#define LJ_NOINLINE __host__ __device__ __noinline__
#else
#define LJ_HD inline
#define LJ_NOINLINE inline
#endif

namespace lj92 {

constexpr int LUT1_BITS = 6;                              // first-level Huffman table (short codes: most symbols)
constexpr int LUT_BITS = 11;                              // second-level table; longer codes: canonical search
constexpr int SUB_BITS = 1024;                            // subsequence length, one per thread
constexpr int SUB_WORDS = SUB_BITS / 32;
constexpr int DEC_THREADS = 256;
constexpr int CHUNK_WORDS = DEC_THREADS * SUB_WORDS;      // 32 KiB of clean stream per thread block
constexpr unsigned CHUNK_BITS = CHUNK_WORDS * 32u;
constexpr int TAIL_WORDS = 4;                             // a symbol (<= 32 bit) may run into the next chunk
constexpr int PHYS_WORDS = CHUNK_WORDS + CHUNK_WORDS / 32 + TAIL_WORDS + 4;
constexpr int UNSTUFF_THREADS = 256;
constexpr int UNSTUFF_CHUNK = UNSTUFF_THREADS * 16;       // payload bytes per unstuff block
constexpr int NCP_MAX = 8;                                // checkpoints per subsequence (decode_span_cp)
constexpr int RING_STRIDE = 67;                           // uint16 per ring row: 64 columns, odd word skew
constexpr int RING_ELEMS = 32 * RING_STRIDE;

enum { ST_OK = 0, ST_HEADER = -1, ST_CORRUPT = -2, ST_INTERNAL = -4 };

struct Tables {                                           // per frame, device memory
    int lw, lh, bits, pred, scan_off, status;
    unsigned clean_bytes, nsym;
    int maxcode[18], mincode[17], valptr[17];
    uint8_t vals[256];
    uint16_t lut1[1 << LUT1_BITS];
    uint16_t lut[1 << LUT_BITS];
};

struct Layout {                                           // byte offsets inside one frame's scratch region
    size_t tables, chunk_cnt, clean, sub_end, sub_cnt, cta_in, cta_cnt, cta_pix, prog, bnd, colsum, tiled, frame_stride, bnd_words;
    unsigned raw_chunks, max_sub, max_cta;
};

struct FrameWork {
    Tables *T;
    unsigned *chunk_cnt;                                  // raw_chunks + 1
    uint8_t *clean;
    uint32_t *sub_end, *sub_cnt;                          // per subsequence: end bit position, symbol count
    uint32_t *cta_in, *cta_cnt, *cta_pix;                 // per decode block: incoming start, symbols, first pixel
    uint32_t *prog;                                       // [0]: strip-group ticket of the prediction kernel
    uint32_t *bnd;                                        // per strip: its last row as tagged words (predict_strip)
    uint32_t *colsum;                                     // per 32-row chunk and 4 columns: packed column sums
    uint16_t *tiled;                                      // differences, then samples, stream ("tiled") order
};

inline size_t lj_align(size_t v) { return (v + 255) / 256 * 256; }

inline Layout make_layout(size_t payload_bytes, size_t npix)
{
    Layout L;
    const size_t bits = payload_bytes * 8;
    L.raw_chunks = (unsigned)((payload_bytes + UNSTUFF_CHUNK - 1) / UNSTUFF_CHUNK);
    L.max_sub = (unsigned)(bits / SUB_BITS + 2);
    L.max_cta = (unsigned)(bits / CHUNK_BITS + 2);
    size_t o = 0;
    L.tables = o;    o += lj_align(sizeof(Tables));
    L.chunk_cnt = o; o += lj_align(sizeof(unsigned) * (L.raw_chunks + 1));
    L.clean = o;     o += lj_align(payload_bytes + 64);
    L.sub_end = o;   o += lj_align(sizeof(uint32_t) * ((size_t)L.max_cta * DEC_THREADS));
    L.sub_cnt = o;   o += lj_align(sizeof(uint32_t) * ((size_t)L.max_cta * DEC_THREADS));
    L.cta_in = o;    o += lj_align(sizeof(uint32_t) * L.max_cta);
    L.cta_cnt = o;   o += lj_align(sizeof(uint32_t) * L.max_cta);
    L.cta_pix = o;   o += lj_align(sizeof(uint32_t) * (L.max_cta + 1));
    L.prog = o;      o += lj_align(sizeof(uint32_t) * 16);
    L.bnd_words = npix / 32 + 65536;                                   // ceil(lh / 32) * lw <= npix / 32 + lw
    L.bnd = o;       o += lj_align(sizeof(uint32_t) * L.bnd_words);
    L.colsum = o;    o += lj_align(npix / 16 + 2 * 65536 + 64);                  // ceil(H / 32) * 2 W bytes
    L.tiled = o;     o += lj_align(sizeof(uint16_t) * npix + 64);
    L.frame_stride = o;
    return L;
}

LJ_HD FrameWork frame_work(char *base, const Layout &L, int frame)
{
    char *p = base + (size_t)frame * L.frame_stride;
    FrameWork F;
    F.T = (Tables *)(p + L.tables);
    F.chunk_cnt = (unsigned *)(p + L.chunk_cnt);
    F.clean = (uint8_t *)(p + L.clean);
    F.sub_end = (uint32_t *)(p + L.sub_end);
    F.sub_cnt = (uint32_t *)(p + L.sub_cnt);
    F.cta_in = (uint32_t *)(p + L.cta_in);
    F.cta_cnt = (uint32_t *)(p + L.cta_cnt);
    F.cta_pix = (uint32_t *)(p + L.cta_pix);
    F.prog = (uint32_t *)(p + L.prog);
    F.bnd = (uint32_t *)(p + L.bnd);
    F.colsum = (uint32_t *)(p + L.colsum);
    F.tiled = (uint16_t *)(p + L.tiled);
    return F;
}

// ------------------------------------------------------------------------------------------------
// 0. headers: marker walk + Huffman tables (serial, one thread per frame)

LJ_HD int parse_headers(const uint8_t *d, int len, Tables &T)
{
    T.lw = T.lh = T.bits = 0; T.pred = -1; T.scan_off = 0; T.clean_bytes = 0; T.nsym = 0;
    for (int i = 0; i < 17; i++) { T.maxcode[i] = -1; T.mincode[i] = 0; T.valptr[i] = 0; }
    T.maxcode[17] = -1;
    if (len < 4 || d[0] != 0xFF || d[1] != 0xD8) return ST_HEADER;
    int ix = 2, scan = -1, have = 0;
    int counts[17];
    for (int i = 0; i < 17; i++) counts[i] = 0;
    while (ix + 4 <= len && scan < 0) {
        if (d[ix] != 0xFF) { ix++; continue; }
        const int marker = d[ix + 1], seg = (d[ix + 2] << 8) | d[ix + 3];
        const uint8_t *s = d + ix + 4;
        if (ix + 2 + seg > len) return ST_HEADER;
        if (marker == 0xC4) {                                                  // DHT, lj92.c:95-271
            int total = 0;
            if (seg < 19) return ST_HEADER;
            for (int i = 1; i <= 16; i++) { counts[i] = s[i]; total += counts[i]; }
            if (total > 256 || 19 + total > seg) return ST_HEADER;
            for (int i = 0; i < total; i++) T.vals[i] = s[17 + i];
            have = 1;
        } else if (marker == 0xC3) {                                           // SOF3, lj92.c:273-280
            if (seg < 7) return ST_HEADER;
            T.bits = s[0];
            T.lh = (s[1] << 8) | s[2];
            T.lw = (s[3] << 8) | s[4];
        } else if (marker == 0xDA) {                                           // SOS, lj92.c:512-519
            const int ncomp = s[0];
            if (3 + 2 * ncomp > seg) return ST_HEADER;
            T.pred = s[1 + 2 * ncomp];
            scan = ix + 2 + seg;
        } else if (marker == 0xD9) return ST_HEADER;
        ix += 2 + seg;
    }
    if (scan < 0 || !have || T.lw <= 0 || T.lh <= 0 || T.pred < 1 || T.pred > 7 || T.bits < 2 || T.bits > 16)
        return ST_HEADER;
    T.scan_off = scan;
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {                                            // canonical code, T.81 Annex C
        T.valptr[l] = k;
        T.mincode[l] = code;
        code += counts[l];
        k += counts[l];
        T.maxcode[l] = counts[l] ? code - 1 : -1;
        code <<= 1;
    }
    return ST_OK;
}

// table entry for the `bits`-bit window `i`: (code length << 8) | (code length + ssss), 0 = a code longer than
// `bits` may start this window, LUT_INVALID = none does (longer code, or an invalid category > 16)
constexpr unsigned LUT_INVALID = 0xFF01u;                 // no code word starts like this: skip one bit

LJ_HD uint16_t lut_entry(const Tables &T, int i, int bits = LUT_BITS)
{
    for (int l = 1; l <= bits; l++) {
        const int code = i >> (bits - l);
        if (T.maxcode[l] >= 0 && code <= T.maxcode[l] && code >= T.mincode[l]) {
            const int v = T.vals[T.valptr[l] + code - T.mincode[l]];
            return v <= 16 ? (uint16_t)((l << 8) | (l + v)) : (uint16_t)LUT_INVALID;
        }
    }
    for (int l = bits + 1; l <= 16; l++)                                       // could a longer code start with i?
        if (T.maxcode[l] >= 0 && (T.mincode[l] >> (l - bits)) <= i && i <= (T.maxcode[l] >> (l - bits))) return 0;
    return (uint16_t)LUT_INVALID;
}

// ------------------------------------------------------------------------------------------------
// 1. unstuff: which of the 16 payload bytes at [q, q+16) belong to the clean stream.
// `pl` is the whole payload (uint32 size + JPEG stream), the entropy-coded segment is [seg0, end).
// A byte is dropped when it follows a 0xFF that was itself kept (lj92.c:356-368: "if (next == 0xff)
// skip one byte", applied serially), i.e. when the run of 0xFF bytes right before it has odd length.

LJ_HD unsigned keep_mask16(const uint8_t *pl, const uint8_t b[16], long long q, long long seg0, long long end)
{
    if (q + 16 <= seg0 || q >= end) return 0;
    int skip = 0;
    if (q > seg0) {
        long long p = q - 1;
        int run = 0;
        while (p >= seg0 && pl[p] == 0xFF) { run++; p--; }
        skip = run & 1;
    }
    if (q >= seg0 && q + 16 <= end) {                                          // common case: no 0xFF among the 16
        uint32_t w[4], any = 0;
        memcpy(w, b, 16);
        for (int i = 0; i < 4; i++) { const uint32_t x = ~w[i]; any |= (x - 0x01010101u) & ~x & 0x80808080u; }
        if (!any) return 0xFFFFu & ~(unsigned)skip;
    }
    unsigned m = 0;
    for (int j = 0; j < 16; j++) {
        const long long p = q + j;
        if (p < seg0) continue;
        if (p >= end) break;
        if (skip) { skip = 0; continue; }
        m |= 1u << j;
        if (b[j] == 0xFF) skip = 1;
    }
    return m;
}

LJ_HD int popc16(unsigned m)
{
#if defined(__CUDA_ARCH__)
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}

// ------------------------------------------------------------------------------------------------
// 2-4. Huffman decode over a 32 KiB chunk held in shared memory

struct DecShared {
    uint32_t words[PHYS_WORDS];            // big-endian stream words, one pad word per 32 (bank skew)
    uint32_t e[2 * DEC_THREADS];           // subsequence ends (local bit positions) / scan scratch
    uint16_t lut1[1 << LUT1_BITS];         // 64 entries = 32 words: one per bank, never a conflict
    uint16_t lut[1 << LUT_BITS];
    uint16_t cp[NCP_MAX * DEC_THREADS];     // checkpoints of decode_span_cp
    int maxcode[18], mincode[17], valptr[17];
    uint8_t vals[256];
};

LJ_HD uint32_t phys(uint32_t k) { return k + (k >> 5); }

// 32 stream bits starting at bit `pos`.  Rows of 32 words are stored 33 apart and the pad word repeats the
// first word of the next row, so the two words straddled by `pos` are always adjacent.
LJ_HD uint32_t window32(const DecShared &S, uint32_t pos)
{
    const uint32_t a = phys(pos >> 5), s = pos & 31;
    const uint32_t w0 = S.words[a], w1 = S.words[a + 1];
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(w1, w0, s);
#else
    return s ? (w0 << s) | (w1 >> (32 - s)) : w0;
#endif
}

// codes longer than LUT_BITS (rare): canonical search, T.81 F.2.2.3.  Kept out of line so that its table
// loads are not hoisted into the symbol loop.
LJ_NOINLINE unsigned decode_symbol_slow(const DecShared &S, uint32_t win)
{
    for (int l = LUT_BITS + 1; l <= 16; l++) {
        const int code = (int)(win >> (32 - l));
        if (S.maxcode[l] >= 0 && code <= S.maxcode[l] && code >= S.mincode[l]) {
            const int v = S.vals[S.valptr[l] + code - S.mincode[l]];
            return v <= 16 ? (unsigned)((l << 8) | (l + v)) : LUT_INVALID;
        }
    }
    return LUT_INVALID;
}

// one symbol = Huffman code of the category ssss + ssss magnitude bits (lj92.c:406-463); returns the
// bits consumed.  An undecodable window (only met off the true parse, or in a corrupt stream)
// advances one bit (entry LUT_INVALID).  Table entries are (code length << 8) | (code length + ssss).
template <bool DIFF>
LJ_HD int decode_symbol(const DecShared &S, uint32_t win, int &diff, int &ok)
{
    unsigned e = S.lut1[win >> (32 - LUT1_BITS)];
    if (!e) {
        e = S.lut[win >> (32 - LUT_BITS)];
        if (!e) e = decode_symbol_slow(S, win);
    }
    const int used = (int)(e & 0xFFu);
    if (DIFF) {
        const int len = (int)(e >> 8), t = used - len;
        diff = 0;
        if (e == LUT_INVALID) { ok = 0; return 1; }
        if (t) {
            int v = (int)((win << len) >> (32 - t));
            if (v < (1 << (t - 1))) v += (int)(0xFFFFFFFFu << t) + 1;         // EXTEND, lj92.c:455-459
            diff = v;
        }
    }
    return used;
}

LJ_HD void decode_span(const DecShared &S, uint32_t pos, uint32_t limit, uint32_t &end, uint32_t &count)
{
    uint32_t n = 0;
    while (pos < limit) {
        int diff, ok = 1;
        pos += decode_symbol<false>(S, window32(S, pos), diff, ok);
        n++;
    }
    end = pos;
    count = n;
}

// Checkpointed form for the synchronisation rounds of dec_body<0>: the subsequence is cut into NCP pieces
// of CP_BITS; when a decode crosses the end of piece j it leaves (symbols in the piece << 5) | (bits past
// the boundary) in S.cp[j][thread].  A later decode from another start stops at the first boundary where
// it lands on the remembered position -- from there on the parse is the one already known -- and
// corrects the symbol count by what changed before that boundary.
constexpr int NCP = 8;
constexpr uint32_t CP_BITS = SUB_BITS / NCP;

template <bool FIRST>
LJ_HD void decode_span_cp(DecShared &S, int tid, uint32_t pos, uint32_t sub_lo, uint32_t limit, uint32_t &end, uint32_t &count)
{
    uint32_t n = 0, nprev = 0, oldcum = 0;
    int j = 0;
    uint32_t nextb = sub_lo + CP_BITS < limit ? sub_lo + CP_BITS : limit;
    while (pos < limit) {
        int diff, ok = 1;
        pos += decode_symbol<false>(S, window32(S, pos), diff, ok);
        n++;
        if (pos >= nextb) {                                                    // at most one boundary per symbol
            uint16_t *slot = &S.cp[j * DEC_THREADS + tid];
            const uint32_t mine = ((n - nprev) << 5) | (pos - nextb);
            if (!FIRST) {
                const uint32_t old = *slot;
                oldcum += old >> 5;
                if ((old & 31u) == pos - nextb) {                              // same position: synchronised
                    *slot = (uint16_t)mine;
                    count = count - oldcum + n;                                // old total - old prefix + new prefix
                    return;                                                    // `end` stays
                }
            }
            *slot = (uint16_t)mine;
            nprev = n;
            j++;
            nextb = nextb + CP_BITS < limit ? nextb + CP_BITS : limit;
        }
    }
    end = pos;
    count = n;
}

// Decode exactly `cnt` symbols from `pos` and store their differences at tiled[idx ...) (indices >= npix
// are dropped: padding after the last pixel).  Stores are 128-bit wherever a group of 8 is aligned --
// the threads of a warp write to 32 different places, so every store is its own L2 transaction.
template <class Ctx>
LJ_HD void decode_write(Ctx &C, const DecShared &S, uint32_t pos, uint32_t cnt, uint16_t *tiled, uint32_t idx, uint32_t npix,
                        uint32_t &end, int &bad)
{
    uint32_t n = 0, head = (8u - (idx & 7u)) & 7u;
    if (head > cnt) head = cnt;
    for (; n < head; n++, idx++) {
        int diff, ok = 1;
        pos += decode_symbol<true>(S, window32(S, pos), diff, ok);
        if (idx < npix) { tiled[idx] = (uint16_t)diff; if (!ok) bad = 1; }
    }
    while (cnt - n >= 8u) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        int notok = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 8; j++) {
            int diff, ok = 1;
            pos += decode_symbol<true>(S, window32(S, pos), diff, ok);
            w[j >> 1] |= (uint32_t)(uint16_t)diff << (16 * (j & 1));
            notok |= (ok ^ 1) << j;
        }
        if (idx < npix) {
            C.store16(tiled + idx, w);                                         // <= 7 samples of slack past npix
            if (notok && (idx + 8 <= npix || (notok & ((1 << (npix - idx)) - 1)))) bad = 1;
        }
        n += 8;
        idx += 8;
    }
    for (; n < cnt; n++, idx++) {
        int diff, ok = 1;
        pos += decode_symbol<true>(S, window32(S, pos), diff, ok);
        if (idx < npix) { tiled[idx] = (uint16_t)diff; if (!ok) bad = 1; }
    }
    end = pos;
}

LJ_HD uint32_t load_be32(const uint8_t *p)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(*reinterpret_cast<const uint32_t *>(p), 0, 0x0123);
#else
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
#endif
}

template <class Ctx>
LJ_HD void load_chunk(Ctx &C, DecShared &S, const FrameWork &F, uint32_t cta, int nwords = CHUNK_WORDS)
{
    const Tables &T = *F.T;
    const size_t nbytes = T.clean_bytes;
    for (int k = C.tid; k < nwords + TAIL_WORDS; k += C.nthr) {
        const size_t off = ((size_t)cta * CHUNK_WORDS + k) * 4;
        uint32_t w = 0;
        if (off + 4 <= nbytes) w = load_be32(F.clean + off);
        else if (off < nbytes)
            for (int j = 0; j < 4 && off + j < nbytes; j++) w |= (uint32_t)F.clean[off + j] << (24 - 8 * j);
        S.words[phys((uint32_t)k)] = w;
        if (k > 0 && (k & 31) == 0) S.words[phys((uint32_t)k) - 1] = w;          // pad word of the row before
    }
    for (int k = C.tid; k < (1 << LUT_BITS); k += C.nthr) S.lut[k] = T.lut[k];
    if (C.tid < (1 << LUT1_BITS)) S.lut1[C.tid] = T.lut1[C.tid];
    for (int k = C.tid; k < 256; k += C.nthr) S.vals[k] = T.vals[k];
    if (C.tid < 18) S.maxcode[C.tid] = T.maxcode[C.tid];
    if (C.tid < 17) { S.mincode[C.tid] = T.mincode[C.tid]; S.valptr[C.tid] = T.valptr[C.tid]; }
    C.sync();
}

// exclusive prefix sum over the block in thread order (Hillis-Steele in shared memory)
template <class Ctx>
LJ_HD uint32_t block_excl_scan(Ctx &C, uint32_t *buf, uint32_t v, uint32_t &total)
{
    const int n = C.nthr, tid = C.tid;
    int in = 0;
    buf[tid] = v;
    C.sync();
    for (int o = 1; o < n; o <<= 1) {
        uint32_t x = buf[in * n + tid];
        if (tid >= o) x += buf[in * n + tid - o];
        buf[(in ^ 1) * n + tid] = x;
        in ^= 1;
        C.sync();
    }
    total = buf[in * n + n - 1];
    const uint32_t r = buf[in * n + tid] - v;
    C.sync();
    return r;
}

// MODE 0: first decode + synchronisation inside the block (start of thread 0 guessed at the chunk start)
// MODE 1: re-synchronise the block after its incoming start changed (no-op when it did not)
// MODE 2: write the differences at their pixel index
template <int MODE, class Ctx>
LJ_HD void dec_body(Ctx &C, DecShared &S, const FrameWork &F, uint32_t cta, uint32_t npix)
{
    const uint32_t total_bits = F.T->clean_bytes * 8u;
    const uint32_t base = cta * CHUNK_BITS;
    if (F.T->status != ST_OK || base >= total_bits) return;                    // uniform over the block
    const int tid = C.tid;
    const uint32_t g = cta * DEC_THREADS + tid, g0 = cta * DEC_THREADS;
    const uint32_t avail = total_bits - base;
    const uint32_t first = (uint32_t)tid * SUB_BITS < avail ? (uint32_t)tid * SUB_BITS : avail;
    const uint32_t lim = (uint32_t)(tid + 1) * SUB_BITS < avail ? (uint32_t)(tid + 1) * SUB_BITS : avail;
    const bool active = first < avail;
    uint32_t inc_g = 0, cin_g = 0;
    if (MODE == 1) {
        if (cta == 0) return;
        // one thread samples the boundary (the predecessor block may be rewriting it in the same launch)
        if (tid == 0) { S.e[0] = F.sub_end[g0 - 1]; S.e[1] = F.cta_in[cta]; }
        C.sync();
        inc_g = S.e[0];
        cin_g = S.e[1];
        C.sync();
        if (inc_g == cin_g) return;                                            // uniform: already consistent
    }
    // a re-synchronisation nearly always ends inside the block's first subsequences: stage only those
    constexpr int FIX_SUBS = 8;
    int loaded = MODE == 1 ? FIX_SUBS : DEC_THREADS;
    load_chunk(C, S, F, cta, loaded * SUB_WORDS);
    int bad = 0;

    if (MODE == 2) {
        const uint32_t start = g == 0 ? 0u : F.sub_end[g - 1];
        if (tid == 0 && cta > 0 && start != F.cta_in[cta]) C.set_status(&F.T->status, ST_INTERNAL);
        const uint32_t cnt = active ? F.sub_cnt[g] : 0u;
        uint32_t tot;
        const uint32_t idx0 = F.cta_pix[cta] + block_excl_scan(C, S.e, cnt, tot);
        if (active) {
            uint32_t e2;
            decode_write(C, S, start - base, cnt, F.tiled, idx0, npix, e2, bad);
            if (e2 + base != F.sub_end[g]) C.set_status(&F.T->status, ST_INTERNAL);
            if (bad) C.set_status(&F.T->status, ST_CORRUPT);
        }
        return;
    }

    uint32_t mystart, myend, n, incoming;
    if (MODE == 0) {
        incoming = 0;
        mystart = first;
        myend = first;
        n = 0;
        decode_span_cp<true>(S, tid, mystart, first, lim, myend, n);
    } else {
        incoming = inc_g - base;
        myend = active ? F.sub_end[g] - base : first;
        n = active ? F.sub_cnt[g] : 0u;
        mystart = tid ? F.sub_end[g - 1] - base : cin_g - base;
    }
    S.e[tid] = myend;
    for (;;) {
        C.sync();
        const uint32_t s = tid ? S.e[tid - 1] : incoming;
        C.sync();
        const bool want = active && s != mystart;
        if (MODE == 1 && loaded < DEC_THREADS) {
            if (C.sync_or(want && tid >= loaded)) {                             // the repair runs past the staged part
                loaded = DEC_THREADS;
                load_chunk(C, S, F, cta, CHUNK_WORDS);
            }
        }
        int changed = 0;
        if (want) {
            mystart = s;
            uint32_t e2 = myend, n2 = n;
            if (MODE == 0) {
                decode_span_cp<false>(S, tid, s, first, lim, e2, n2);
            } else {
                decode_span(S, s, lim, e2, n2);
            }
            changed = e2 != myend;
            myend = e2;
            n = n2;
            S.e[tid] = e2;
        }
        if (!C.sync_or(changed)) break;
    }
    if (active) { F.sub_end[g] = base + myend; F.sub_cnt[g] = n; }
    uint32_t tot;
    block_excl_scan(C, S.e, active ? n : 0u, tot);
    if (tid == 0) { F.cta_cnt[cta] = tot; F.cta_in[cta] = base + incoming; }
}

// One block per frame, after the parallel passes: walk the block boundaries in stream order and
// re-synchronise every block whose incoming start is not the end its predecessor reports.  Every fix
// is final for that block (its predecessor is), so this converges for any stream.  Then prefix-sum
// the per-block symbol counts.
template <class Ctx>
LJ_HD void resolve_body(Ctx &C, DecShared &S, const FrameWork &F, uint32_t npix)
{
    if (F.T->status != ST_OK) return;
    const uint32_t total_bits = F.T->clean_bytes * 8u;
    const uint32_t ncta = (uint32_t)(((unsigned long long)total_bits + CHUNK_BITS - 1) / CHUNK_BITS);
    const int tid = C.tid, n = C.nthr;
    uint32_t lo = 1;
    for (;;) {
        uint32_t mine = 0xFFFFFFFFu;
        for (uint32_t b = lo + tid; b < ncta; b += n)
            if (F.sub_end[b * DEC_THREADS - 1] != F.cta_in[b]) { mine = b; break; }
        S.e[tid] = mine;
        C.sync();
        for (int o = n >> 1; o > 0; o >>= 1) {
            if (tid < o && S.e[tid + o] < S.e[tid]) S.e[tid] = S.e[tid + o];
            C.sync();
        }
        const uint32_t firstbad = S.e[0];
        C.sync();
        if (firstbad == 0xFFFFFFFFu) break;
        dec_body<1>(C, S, F, firstbad, npix);
        C.sync();
        lo = firstbad + 1;
    }
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < ncta; b0 += n) {
        const uint32_t b = b0 + tid;
        const uint32_t v = b < ncta ? F.cta_cnt[b] : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(C, S.e, v, tot);
        if (b < ncta) F.cta_pix[b] = carry + ex;
        carry += tot;
    }
    if (tid == 0) {
        F.cta_pix[ncta] = carry;
        F.T->nsym = carry;
        if (carry < npix) C.set_status(&F.T->status, ST_CORRUPT);              // stream ends before the image does
    }
}

// ------------------------------------------------------------------------------------------------
// 5. prediction (lj92.c:408-593) as a skewed wavefront, in place on the `tiled` buffer

LJ_HD int predict(int pred, int a, int b, int c)
{
    switch (pred) {
    case 1: return a;
    case 2: return b;
    case 3: return c;
    case 4: return a + b - c;
    case 5: return a + ((b - c) >> 1);
    case 6: return b + ((a - c) >> 1);                                         // lj92.c:488
    default: return (a + b) >> 1;
    }
}

// Predictor 6 separates.  With U[r][c] = x[r][c] - x[r-1][c] (sample minus the sample above), the
// predictor  x[r-1][c] + ((x[r][c-1] - x[r-1][c-1]) >> 1) + d  (lj92.c:488) reads
//     U[r][c] = (U[r][c-1] >> 1) + d[r][c],   U[r][-1] = 0            -- a recurrence along the row only,
//     x[r][c] = x[r-1][c] + U[r][c]                                   -- a prefix sum down the column,
// and row 0 (predicted from the left, lj92.c:441-456) is U[0][c] = U[0][c-1] + d with U[0][-1] = 2^(bits-1),
// x[-1][c] = 0.  Exact in integers while samples fit 16 bits and differences fit int16 (bits <= 15); the
// column sums are taken modulo 2^16 like the reference's uint16 rows.  Streams with another predictor,
// 16-bit samples or a raster narrower/wider than the frame take the wavefront (predict_body).
constexpr int CH_ROWS = 32;                               // rows per column-sum chunk

LJ_HD bool separable(const Tables &T, int W) { return T.pred == 6 && T.bits <= 15 && T.lw == W; }
LJ_HD int row_step(bool row0, int U, int d) { return row0 ? U + d : (U >> 1) + d; }

constexpr uint32_t BND_READY = 0x10000u;                 // tag bit of a boundary word (sample | tag)

// The dec_body<2> launch clears the boundary words (and the strip ticket) of its frame; `part` of
// `nparts` blocks each clear a slice.
template <class Ctx>
LJ_HD void clear_boundary(Ctx &C, const FrameWork &F, const Layout &L, uint32_t part, uint32_t nparts)
{
    const size_t n = L.bnd_words, per = (n + nparts - 1) / nparts;
    const size_t lo = (size_t)part * per, hi = lo + per < n ? lo + per : n;
    for (size_t i = lo + C.tid; i < hi; i += C.nthr) F.bnd[i] = 0;
    if (part == 0 && C.tid == 0) F.prog[0] = 0;
}

// One 32-row strip.  Lane l owns row 32 s + l and, at step t, column t - l, so "above" is what lane
// l - 1 produced one step earlier (one shuffle); its row of the current and the next 32-column block
// sits in its private row of `ring` (64 columns), refilled / flushed one block at a time.  Lane 0's
// "above" is the last row of strip s - 1, which that strip's lane 31 publishes sample by sample as
// tagged words in F.bnd (sample | BND_READY): one aligned 32-bit store carries data and flag, so no
// fence is needed and the strip below follows two blocks behind.
template <int PRED, class Ctx>
LJ_HD void predict_strip(Ctx &C, uint16_t *ring_warp, const FrameWork &F, int s)
{
    const Tables &T = *F.T;
    const int lw = T.lw, lh = T.lh, pred = T.pred, first_px = 1 << (T.bits - 1);
    const int nstrips = (lh + 31) / 32, nblk = (lw + 31) / 32;
    const int lane = C.tid & 31;
    uint16_t *ring = ring_warp + lane * RING_STRIDE;
    const int r = s * 32 + lane;
    const bool rowok = r < lh, row0 = r == 0, vec = (lw % 32) == 0;
    const bool publish = lane == 31 && s + 1 < nstrips;
    uint16_t *row = F.tiled + (size_t)r * lw;
    const uint32_t *bnd_in = F.bnd + (size_t)(s > 0 ? s - 1 : 0) * lw;
    uint32_t *bnd_out = F.bnd + (size_t)s * lw;
    uint32_t nx[16];
    uint32_t abv = 0, abw = BND_READY, left = 0, prevb = 0;
    int abcol = -1;                                       // column abw was (or has to be) loaded from, -1: none

    // prefetch block kk: this lane's 32 differences, and one boundary word of the row above
#define LJ_PREFETCH(kk)                                                                                     \
    do {                                                                                                    \
        const int c0_ = (kk) * 32;                                                                          \
        if (rowok && vec) C.load64(row + c0_, nx);                                                          \
        else                                                                                                \
            for (int m = 0; m < 16; m++) {                                                                  \
                const int ca = c0_ + 2 * m;                                                                 \
                const uint32_t v0 = (rowok && ca < lw) ? row[ca] : 0u, v1 = (rowok && ca + 1 < lw) ? row[ca + 1] : 0u; \
                nx[m] = v0 | (v1 << 16);                                                                    \
            }                                                                                               \
        abcol = (s > 0 && c0_ + lane < lw) ? c0_ + lane : -1;                                               \
        abw = abcol >= 0 ? C.load_volatile32(bnd_in + abcol) : BND_READY;                                   \
    } while (0)

#define LJ_FLUSH(kk)                                                                                        \
    do {                                                                                                    \
        const int c0_ = (kk) * 32, sl_ = ((kk) & 1) * 32;                                                   \
        if (rowok) {                                                                                        \
            if (vec) {                                                                                      \
                uint32_t o_[16];                                                                            \
                for (int m = 0; m < 16; m++) o_[m] = ring[sl_ + 2 * m] | ((uint32_t)ring[sl_ + 2 * m + 1] << 16); \
                C.store64(row + c0_, o_);                                                                   \
            } else                                                                                          \
                for (int m = 0; m < 32; m++) if (c0_ + m < lw) row[c0_ + m] = ring[sl_ + m];                \
        }                                                                                                   \
    } while (0)

    LJ_PREFETCH(0);
    for (int k = 0; k <= nblk; k++) {
        const int sl = (k & 1) * 32;
        if (k >= 2) LJ_FLUSH(k - 2);
        if (k < nblk) {
            for (int m = 0; m < 16; m++) { ring[sl + 2 * m] = (uint16_t)nx[m]; ring[sl + 2 * m + 1] = (uint16_t)(nx[m] >> 16); }
            while (!C.all((abw & BND_READY) != 0)) {                          // the strip above is not there yet
                if (!(abw & BND_READY)) { C.pause(); abw = C.load_volatile32(bnd_in + abcol); }
            }
            abv = abw & 0xFFFFu;
        }
        if (k + 1 < nblk) LJ_PREFETCH(k + 1);
        const int cbase = k * 32 - lane;
        for (int tt = 0; tt < 32; tt++) {
            const uint32_t b_up = C.shfl_up(left, 1), b0 = C.shfl(abv, tt);
            const uint32_t bv = lane ? b_up : b0;
            const int c = cbase + tt;
            const bool act = rowok && (unsigned)c < (unsigned)lw;
            uint16_t *cell = ring + (c & 63);
            const int d = (int16_t)*cell;
            int px = PRED == 6 ? (int)bv + (((int)left - (int)prevb) >> 1) : predict(pred, (int)left, (int)bv, (int)prevb);
            if (c == 0) px = (int)bv;
            if (row0) px = c == 0 ? first_px : (int)left;
            const uint32_t nl = (uint32_t)(px + d) & 0xFFFFu;
            if (act) {
                *cell = (uint16_t)nl;
                left = nl;
                if (publish) C.store_volatile32(bnd_out + c, nl | BND_READY);
            }
            prevb = bv;
        }
    }
    LJ_FLUSH(nblk - 1);
#undef LJ_PREFETCH
#undef LJ_FLUSH
}

// A block takes groups of `nwarps` consecutive strips by ticket (F.prog[0]) until none are left.  A
// strip only ever waits for the strip above it, which belongs to the same or an earlier ticket, i.e. to
// a block that is already running: no deadlock, whatever the grid size and dispatch order.
template <class Ctx>
LJ_HD void predict_body(Ctx &C, uint16_t *ring_all, uint32_t *s_ticket, const FrameWork &F, bool skip_separable, int W)
{
    const Tables &T = *F.T;
    if (T.status != ST_OK || (skip_separable && separable(T, W))) return;
    const int nstrips = (T.lh + 31) / 32;
    const int warp = C.tid >> 5, nwarps = C.nthr >> 5;
    for (;;) {
        if (C.tid == 0) *s_ticket = C.atomic_add(&F.prog[0], 1u);
        C.sync();
        const uint32_t part = *s_ticket;
        C.sync();
        if ((long long)part * nwarps >= nstrips) break;
        const int s = (int)part * nwarps + warp;
        if (s < nstrips) {
            if (T.pred == 6) predict_strip<6>(C, ring_all + warp * RING_ELEMS, F, s);
            else predict_strip<0>(C, ring_all + warp * RING_ELEMS, F, s);
        }
    }
}

}  // namespace lj92
