// lj92.cu -- lossless-JPEG ("LJ92") VIDF payload -> 16-bit frame, de-interleaved in the same pass.
//
// Replaces reference lj92.c:650-702 (lj92_open / lj92_decode: marker parse lj92.c:83-280, Huffman
// bit reader with 0xFF00 stuffing lj92.c:344-406, predictors lj92.c:408-593) and the quadrant
// de-interleave loop of main.c:656-668.
//
// A lossless-JPEG scan has no restart markers, so decoding one frame is a serial walk over its bit
// stream (position of symbol n+1 depends on symbol n; pixel n+1 is predicted from pixel n).  As the
// north star prescribes, the parallelism is ACROSS frames: one warp per frame.  Lane 0 walks the
// stream 32 pixels at a time out of shared memory; all 32 lanes stage the next 512 stream bytes and
// the row above into shared memory before each round and scatter the 32 finished pixels to their
// de-interleaved positions after it, so global traffic stays coalesced and off the serial chain.
// This stage is latency-bound (report: issue-slot use, not HBM).
#include "kernels.cuh"

namespace {

constexpr int LJ_LUT_BITS = 10;
constexpr int LJ_WARPS = 4;          // frames per CTA
constexpr int LJ_STAGE = 512;        // stream bytes staged per round (32 px * <=32 bit, stuffing-safe)

struct alignas(16) WarpState {
    uint16_t lut[1 << LJ_LUT_BITS];  // (ssss << 8) | code length, 0 = longer than LJ_LUT_BITS
    alignas(16) uint8_t stage[LJ_STAGE + 16];
    uint16_t above[34];
    uint16_t cur[32];
    int maxcode[18], mincode[17], valptr[17];
    uint8_t vals[256];
    int hdr[8];                      // w, h, bits, pred, scan offset, status
};

__device__ __forceinline__ size_t untiled_index(unsigned c, int W, int H)
{
    const unsigned y = c / W, x = c - y * W;                       // main.c:656-668
    const unsigned dy = (2 * y) % H + (2 * y) / H, dx = (2 * x) % W + (2 * x) / W;
    return (size_t)dy * W + dx;
}

// marker walk + canonical Huffman tables; lane 0 only
__device__ int parse_headers(const uint8_t *d, int len, WarpState &S)
{
    if (len < 4 || d[0] != 0xFF || d[1] != 0xD8) return -1;
    int ix = 2, scan = -1, have = 0;
    int counts[17];
    for (int i = 0; i < 17; i++) counts[i] = 0;
    S.hdr[0] = S.hdr[1] = S.hdr[2] = 0; S.hdr[3] = -1;
    while (ix + 4 <= len && scan < 0) {
        if (d[ix] != 0xFF) { ix++; continue; }
        const int marker = d[ix + 1], seg = (d[ix + 2] << 8) | d[ix + 3];
        const uint8_t *s = d + ix + 4;
        if (ix + 2 + seg > len) return -1;
        if (marker == 0xC4) {
            int total = 0;
            for (int i = 1; i <= 16; i++) { counts[i] = s[i]; total += counts[i]; }
            if (total > 256 || 17 + total > seg) return -1;
            for (int i = 0; i < total; i++) S.vals[i] = s[17 + i];
            have = 1;
        } else if (marker == 0xC3) {
            S.hdr[2] = s[0];
            S.hdr[1] = (s[1] << 8) | s[2];
            S.hdr[0] = (s[3] << 8) | s[4];
        } else if (marker == 0xDA) {
            const int ncomp = s[0];
            S.hdr[3] = s[1 + 2 * ncomp];
            scan = ix + 2 + seg;
        } else if (marker == 0xD9) return -1;
        ix += 2 + seg;
    }
    if (scan < 0 || !have || S.hdr[0] <= 0 || S.hdr[1] <= 0 || S.hdr[3] < 1 || S.hdr[3] > 7) return -1;
    S.hdr[4] = scan;
    for (int i = 0; i < (1 << LJ_LUT_BITS); i++) S.lut[i] = 0;
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
        S.valptr[l] = k;
        S.mincode[l] = code;
        for (int c = 0; c < counts[l]; c++, k++, code++) {
            if (l <= LJ_LUT_BITS) {
                const int lo = code << (LJ_LUT_BITS - l), n = 1 << (LJ_LUT_BITS - l);
                const uint16_t e = (uint16_t)((S.vals[k] << 8) | l);
                for (int j = 0; j < n; j++) S.lut[lo + j] = e;
            }
        }
        S.maxcode[l] = counts[l] ? code - 1 : -1;
        code <<= 1;
    }
    return 0;
}

__global__ void __launch_bounds__(LJ_WARPS * 32)
lj92_decode_kernel(const uint8_t *__restrict__ payload_base, size_t payload_stride, size_t payload_bytes,
                   uint16_t *__restrict__ out_base, size_t out_stride_px, int W, int H, int nframes,
                   int *__restrict__ status)
{
    __shared__ WarpState WS[LJ_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int frame = blockIdx.x * LJ_WARPS + wid;
    if (frame >= nframes) return;
    WarpState &S = WS[wid];
    // payload = uint32 stored size, then the JPEG stream (main.c:626-629)
    const uint8_t *data = payload_base + (size_t)frame * payload_stride + 4;
    const int len = (int)min(payload_bytes - 4, (size_t)0x7FFFFFF0);
    uint16_t *out = out_base + (size_t)frame * out_stride_px;

    if (lane == 0) S.hdr[5] = parse_headers(data, len, S);
    __syncwarp();
    const int lw = S.hdr[0], lh = S.hdr[1], depth = S.hdr[2], pred = S.hdr[3];
    if (S.hdr[5] != 0 || (long long)lw * lh != (long long)W * H) {
        if (lane == 0) status[frame] = -1;
        return;
    }
    const unsigned npix = (unsigned)lw * lh;

    // lane 0's serial state
    unsigned long long acc = 0;
    int nb = 0;
    long long spos = S.hdr[4];          // stream byte position of the next unread byte
    int left = 0;
    int bad = 0;

    for (unsigned c0 = 0; c0 < npix;) {
        const unsigned col0 = c0 % lw, row = c0 / lw;
        const int n = (int)min(32u, (unsigned)lw - col0);
        // --- stage stream bytes and the row above (all lanes)
        const long long sbase = spos & ~15ll;
        {
            const long long off = sbase + 16 * lane;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (off + 16 <= len) {
                if ((((uintptr_t)(data + off)) & 15) == 0) v = __ldg(reinterpret_cast<const uint4 *>(data + off));
                else {
                    uint8_t *b = reinterpret_cast<uint8_t *>(&v);
                    for (int i = 0; i < 16; i++) b[i] = __ldg(data + off + i);
                }
            } else {
                uint8_t *b = reinterpret_cast<uint8_t *>(&v);
                for (int i = 0; i < 16; i++) b[i] = (off + i < len) ? __ldg(data + off + i) : (uint8_t)0;
            }
            *reinterpret_cast<uint4 *>(&S.stage[16 * lane]) = v;
        }
        if (row > 0) {
            // above[j] = pixel (row-1, col0 + j - 1), j = 0..32
            const long long cj = (long long)c0 - lw + lane - 1;
            if (col0 + lane >= 1 && col0 + lane - 1 < (unsigned)lw) S.above[lane] = __ldcg(out + untiled_index((unsigned)cj, W, H));
            if (lane == 31 && col0 + 31 < (unsigned)lw) S.above[32] = __ldcg(out + untiled_index((unsigned)(cj + 1), W, H));
        }
        __syncwarp();

        // --- lane 0: decode n pixels
        if (lane == 0 && !bad) {
            int sp = (int)(spos - sbase);
            for (int i = 0; i < n; i++) {
                while (nb <= 56) {
                    const unsigned byte = S.stage[sp++];
                    acc = (acc << 8) | byte;
                    nb += 8;
                    if (byte == 0xFF) sp++;                              // stuffed zero (lj92.c:356-368)
                }
                const unsigned peek = (unsigned)(acc >> (nb - LJ_LUT_BITS)) & ((1u << LJ_LUT_BITS) - 1);
                const unsigned e = S.lut[peek];
                int t, l;
                if (e) { l = e & 0xFF; t = e >> 8; }
                else {
                    t = -1;
                    for (l = LJ_LUT_BITS + 1; l <= 16; l++) {
                        const int code = (int)((acc >> (nb - l)) & ((1u << l) - 1));
                        if (S.maxcode[l] >= 0 && code <= S.maxcode[l] && code >= S.mincode[l]) { t = S.vals[S.valptr[l] + code - S.mincode[l]]; break; }
                    }
                    if (t < 0) { bad = 1; break; }
                }
                nb -= l;
                int diff = 0;
                if (t) {
                    nb -= t;
                    diff = (int)((acc >> nb) & ((1u << t) - 1));
                    if (diff < (1 << (t - 1))) diff += (int)(0xFFFFFFFFu << t) + 1;
                }
                const unsigned col = col0 + i;
                int px;
                if (row == 0) px = col == 0 ? (1 << (depth - 1)) : left;
                else if (col == 0) px = S.above[1];
                else {
                    const int a = left, b = S.above[i + 1], cc = S.above[i];
                    switch (pred) {
                    case 1: px = a; break;
                    case 2: px = b; break;
                    case 3: px = cc; break;
                    case 4: px = a + b - cc; break;
                    case 5: px = a + ((b - cc) >> 1); break;
                    case 6: px = b + ((a - cc) >> 1); break;             // lj92.c:488
                    default: px = (a + b) >> 1; break;
                    }
                }
                left = px + diff;
                S.cur[i] = (uint16_t)left;
                left = (uint16_t)left;        // the reference re-reads 16-bit row values for the next row only; keep both in range
            }
            // bytes still buffered in acc were consumed from the stage: keep the byte position exact
            spos = sbase + sp;
        }
        spos = __shfl_sync(0xFFFFFFFFu, spos, 0);
        bad = __shfl_sync(0xFFFFFFFFu, bad, 0);
        __syncwarp();
        if (bad) break;
        // --- scatter the finished pixels to de-interleaved positions (all lanes)
        if (lane < n) out[untiled_index(c0 + lane, W, H)] = S.cur[lane];
        __syncwarp();
        c0 += n;
    }
    if (lane == 0) status[frame] = bad ? -2 : 0;
}

}  // namespace

int launch_lj92_decode(const void *d_payload, size_t payload_stride, size_t payload_bytes, uint16_t *d_out,
                       size_t out_stride_px, int w, int h, int nframes, int *d_status, cudaStream_t st)
{
    if (payload_bytes < 8) return MLVB_ERR_ARG;
    lj92_decode_kernel<<<ceil_div(nframes, LJ_WARPS), LJ_WARPS * 32, 0, st>>>((const uint8_t *)d_payload, payload_stride,
                                                                             payload_bytes, d_out, out_stride_px, w, h,
                                                                             nframes, d_status);
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}
