// kernels.cuh -- internal launcher interface between the C-ABI layer (abi.cu) and the kernels.
// Every launcher enqueues on the given stream and returns MLVB_OK or a negative MLVB_ERR_* code;
// none of them synchronises.
#pragma once
#include "common.cuh"

struct StripeCoef {          // reference stripes.h:30-36 (struct stripes_correction), minus the list plumbing
    int needed;
    int coef[8];
};

struct PixelXY { int x, y; };  // reference cs.c:170-174 (struct focus_pixel)

// ---- unpack.cu ----
size_t mlvb_packed_bytes(uint32_t npix, int bpp);
int launch_unpack(const void *d_in, size_t in_stride_bytes, size_t in_bytes, uint16_t *d_out, size_t out_stride_px,
                  uint32_t npix, int bpp, int nframes, cudaStream_t st);
int launch_unpack_range(const uint16_t *d_in, uint16_t *d_out, uint32_t first_px, uint32_t npix, int bpp,
                        cudaStream_t st);

// ---- chroma.cu ----
int launch_chroma_smooth_u16(const uint16_t *d_in, uint16_t *d_out, int w, int h, size_t frame_stride, int nframes,
                             int black, int method, const EvLuts &luts, const StripeCoef *stripes, int white,
                             cudaStream_t st);
int launch_chroma_smooth_u32(const uint32_t *d_in, uint32_t *d_out, int w, int h, int method, const int *d_raw2ev,
                             const int *d_ev2raw, cudaStream_t st, const uint8_t *d_dead_flags = nullptr, int dead_mask = 0);
int launch_stripes_apply(uint16_t *d_img, int w, size_t npix, size_t frame_stride, int nframes, int black, int white,
                         const StripeCoef *sc, cudaStream_t st);

// ---- pixfix.cu ----
int badpix_detect_scratch_bytes(int w, int h, size_t *flag_bytes, size_t *count_bytes);   // returns #CTAs
int launch_badpix_detect_count(const uint16_t *d_img, int w, int h, int black, int aggressive, const EvLuts &luts,
                               uint8_t *d_flags, unsigned long long *d_counts, cudaStream_t st);
int launch_badpix_detect_scatter(const uint8_t *d_flags, const unsigned long long *d_offsets, int w, int h, int crop_x,
                                 int crop_y, PixelXY *d_list, cudaStream_t st);
int launch_pixel_fix(uint16_t *d_img, int w, int h, size_t frame_stride, int nframes, int black, int crop_x, int crop_y,
                     int dual_iso, int edge_rules, const PixelXY *d_list_by_level, const unsigned *d_level_start,
                     const unsigned *h_level_start, unsigned nlevels, const EvLuts &luts, cudaStream_t st);

// horizontal-only (dual ISO) form: independent row segments, one thread each
int launch_pixel_fix_rows(uint16_t *d_img, int w, int h, size_t frame_stride, int nframes, int black, int crop_x, int crop_y,
                          int edge_rules, const PixelXY *d_list_by_row, const unsigned *d_seg_start, unsigned nseg,
                          const unsigned *d_long_rows, unsigned nlong, const EvLuts &luts, int ev2raw_octaves_ok, int sm_count,
                          cudaStream_t st);

// ---- stripes.cu ----
int stripes_blocks_per_row(int w);
int stripes_count_ctas(int w, int h);
int launch_stripes_count(const uint16_t *d_img, int w, int h, int black, int white, unsigned long long *d_counts,
                         cudaStream_t st);
int launch_stripes_hist(const uint16_t *d_img, int w, int h, int black, int white, const unsigned long long *d_offsets,
                        const uint16_t *d_dither, unsigned *d_hist, unsigned *d_num, int *d_median_bin, cudaStream_t st);

// ---- lj92.cu ----
size_t lj92_scratch_bytes(size_t payload_bytes, size_t npix, int nframes);
// returns the number of kernels launched (> 0) or a negative MLVB_ERR_* code
int launch_lj92_decode(const void *d_payload, size_t payload_stride, size_t payload_bytes, uint16_t *d_out,
                       size_t out_stride_px, int w, int h, int nframes, int *d_status, void *d_scratch,
                       size_t scratch_bytes, cudaStream_t st);

// ---- lzma_dec.cu (host code): raw LZMA1 stream -> bytes; *dst_len in = capacity, out = bytes produced ----
int mlvb_lzma_decode(uint8_t *dst, size_t *dst_len, const uint8_t *src, size_t src_len, const uint8_t props[5]);

// ---- patternnoise.cu ----
size_t pattern_noise_scratch_bytes(int w, int h);
int launch_pattern_noise(int16_t *d_raw, int w, int h, int white, void *d_scratch, cudaStream_t st);
