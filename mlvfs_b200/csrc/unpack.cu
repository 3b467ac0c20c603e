// unpack.cu -- packed 8/10/12/14-bit MLV payload -> 16-bit little-endian DNG samples.
//
// Replaces reference mlvfs/dng.c:813-872 (dng_get_image_data / dng_get_image_data_inline).
// Bit layout (reference raw.h:41-79): the payload is a stream of 16-bit LE words read
// most-significant-bit first; pixel i is bits [i*bpp, (i+1)*bpp).  Eight pixels therefore occupy
// exactly `bpp` bytes, which is the unit ("group") the fast kernel works in.
//
// HBM-bound: 1.75 B/px in + 2 B/px out at 14 bit.  The packed stream is staged through shared
// memory with coalesced 128-bit loads; each thread then extracts one 8-pixel group and issues one
// coalesced 128-bit store.
#include "kernels.cuh"

// default cache policy for the stream and the frame (measured on B200: 783 k frames/s against 766 k with the
// streaming hints __ldcs / __stcs, -DUNPACK_STREAMING_HINTS)
#ifdef UNPACK_STREAMING_HINTS
#define UNPACK_ST(p, v) __stcs(p, v)
#define UNPACK_LD(p) __ldcs(p)
#else
#define UNPACK_ST(p, v) (*(p) = (v))
#define UNPACK_LD(p) (*(p))
#endif

namespace {

constexpr int UNPACK_THREADS = 256;
constexpr int UNPACK_GROUPS = 1024;  // 8-pixel groups per CTA: 8192 px, 14 KiB in, 16 KiB out at 14 bit

template <int BPP>
__device__ __forceinline__ uint32_t extract_px(const uint16_t *w, int j)
{
    const int bit = j * BPP, k = bit >> 4, s = bit & 15;
    uint32_t pair = (uint32_t)w[k] << 16;
    if (s + BPP > 16) pair |= w[k + 1];
    return (pair >> (32 - BPP - s)) & ((1u << BPP) - 1);
}

template <int BPP>
__global__ void __launch_bounds__(UNPACK_THREADS)
unpack_groups_kernel(const uint8_t *__restrict__ in, size_t in_stride, size_t in_bytes,
                     uint16_t *__restrict__ out, size_t out_stride_px, uint32_t npix)
{
    constexpr int WORDS = BPP / 2;                       // 16-bit words per group
    constexpr int CHUNK_BYTES = UNPACK_GROUPS * BPP;     // multiple of 16
    __shared__ __align__(16) uint16_t stage[CHUNK_BYTES / 2];

    const uint32_t chunk = blockIdx.x;
    const uint8_t *src = in + (size_t)blockIdx.y * in_stride + (size_t)chunk * CHUNK_BYTES;
    uint16_t *dst = out + (size_t)blockIdx.y * out_stride_px + (size_t)chunk * UNPACK_GROUPS * 8;

    const size_t chunk_off = (size_t)chunk * CHUNK_BYTES;
    const uint32_t avail = chunk_off < in_bytes ? (uint32_t)min((size_t)CHUNK_BYTES, in_bytes - chunk_off) : 0u;

    // stage: 128-bit coalesced loads while a full vector is inside the payload, words for the tail
    for (uint32_t v = threadIdx.x; v < CHUNK_BYTES / 16; v += UNPACK_THREADS) {
        const uint32_t off = v * 16;
        if (off + 16 <= avail) {
            reinterpret_cast<uint4 *>(stage)[v] = UNPACK_LD(reinterpret_cast<const uint4 *>(src + off));
        } else {
            for (uint32_t b = off; b < off + 16; b += 2)
                stage[b / 2] = (b + 2 <= avail) ? *reinterpret_cast<const uint16_t *>(src + b) : (uint16_t)0;
        }
    }
    __syncthreads();

    const uint32_t groups_total = npix / 8;
    const uint32_t first_group = chunk * UNPACK_GROUPS;
    const uint32_t ngroups = min((uint32_t)UNPACK_GROUPS, groups_total - min(groups_total, first_group));

#pragma unroll
    for (int r = 0; r < UNPACK_GROUPS / UNPACK_THREADS; r++) {
        const uint32_t g = threadIdx.x + r * UNPACK_THREADS;
        if (g < ngroups) {
            uint16_t w[WORDS + 1];
#pragma unroll
            for (int k = 0; k < WORDS; k++) w[k] = stage[g * WORDS + k];
            w[WORDS] = 0;
            uint32_t p[8];
#pragma unroll
            for (int j = 0; j < 8; j++) p[j] = extract_px<BPP>(w, j);
            uint4 o;
            o.x = p[0] | (p[1] << 16);
            o.y = p[2] | (p[3] << 16);
            o.z = p[4] | (p[5] << 16);
            o.w = p[6] | (p[7] << 16);
            UNPACK_ST(reinterpret_cast<uint4 *>(dst) + g, o);
        }
    }

    // up to 7 trailing pixels of the frame (npix % 8) live in the chunk that holds group `groups_total`
    const uint32_t rem = npix & 7;
    if (rem && groups_total >= first_group && groups_total < first_group + UNPACK_GROUPS && threadIdx.x < rem) {
        const uint32_t g = groups_total - first_group;
        const uint32_t bit = threadIdx.x * BPP, k = bit >> 4, s = bit & 15;
        uint32_t pair = (uint32_t)stage[g * WORDS + k] << 16;
        if (s + BPP > 16) pair |= stage[g * WORDS + k + 1];
        dst[g * 8 + threadIdx.x] = (uint16_t)((pair >> (32 - BPP - s)) & ((1u << BPP) - 1));
    }
}

// Any bpp (1..16), any starting pixel: one pixel per thread.  `in` points at the word holding
// pixel `first_px` (reference dng.c:816-822 pointer arithmetic).
__global__ void unpack_generic_kernel(const uint16_t *__restrict__ in, uint16_t *__restrict__ out,
                                      uint32_t first_px, uint32_t npix, int bpp)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npix) return;
    const uint32_t first_word = first_px * (uint32_t)bpp / 16;
    const uint32_t bit = (first_px + j) * (uint32_t)bpp;
    const uint32_t k = bit / 16 - first_word, s = bit % 16;
    uint32_t pair = (uint32_t)in[k] << 16;
    if (s + bpp > 16) pair |= in[k + 1];
    out[j] = (uint16_t)((pair >> (32 - bpp - s)) & ((1u << bpp) - 1));
}

template <int BPP>
void launch_groups(const void *in, size_t in_stride, size_t in_bytes, uint16_t *out, size_t out_stride_px,
                   uint32_t npix, int nframes, cudaStream_t st)
{
    const uint32_t groups = (npix + 7) / 8;
    dim3 grid((groups + UNPACK_GROUPS - 1) / UNPACK_GROUPS, nframes);
    unpack_groups_kernel<BPP><<<grid, UNPACK_THREADS, 0, st>>>((const uint8_t *)in, in_stride, in_bytes, out,
                                                               out_stride_px, npix);
}

}  // namespace

size_t mlvb_packed_bytes(uint32_t npix, int bpp) { return (((size_t)npix * bpp + 15) / 16) * 2; }

int launch_unpack(const void *d_in, size_t in_stride_bytes, size_t in_bytes, uint16_t *d_out, size_t out_stride_px,
                  uint32_t npix, int bpp, int nframes, cudaStream_t st)
{
    if (npix == 0 || nframes == 0) return MLVB_OK;
    const bool aligned = ((uintptr_t)d_in % 16 == 0) && ((uintptr_t)d_out % 16 == 0) && (in_stride_bytes % 16 == 0) &&
                         (out_stride_px % 8 == 0);
    if (aligned && (bpp == 14 || bpp == 12 || bpp == 10)) {
        if (bpp == 14) launch_groups<14>(d_in, in_stride_bytes, in_bytes, d_out, out_stride_px, npix, nframes, st);
        if (bpp == 12) launch_groups<12>(d_in, in_stride_bytes, in_bytes, d_out, out_stride_px, npix, nframes, st);
        if (bpp == 10) launch_groups<10>(d_in, in_stride_bytes, in_bytes, d_out, out_stride_px, npix, nframes, st);
    } else {
        if (bpp < 1 || bpp > 16 || in_stride_bytes % 2) return MLVB_ERR_ARG;
        for (int f = 0; f < nframes; f++)
            unpack_generic_kernel<<<ceil_div(npix, 256), 256, 0, st>>>(
                (const uint16_t *)((const uint8_t *)d_in + f * in_stride_bytes), d_out + f * out_stride_px, 0, npix, bpp);
    }
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}

int launch_unpack_range(const uint16_t *d_in, uint16_t *d_out, uint32_t first_px, uint32_t npix, int bpp, cudaStream_t st)
{
    if (npix == 0) return MLVB_OK;
    if (bpp < 1 || bpp > 16) return MLVB_ERR_ARG;
    unpack_generic_kernel<<<ceil_div(npix, 256), 256, 0, st>>>(d_in, d_out, first_px, npix, bpp);
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}
