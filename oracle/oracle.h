/*
 * oracle/oracle.h -- CPU restatement of the MLVFS per-frame raw path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mlvfs_b200/ (the product) may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker.
 *
 * Parity pin: the reference ships no tests, goldens or KATs for this path, so
 * the restatement is pinned against the UNMODIFIED reference compiled into oracle/_ref
 * (tests/test_oracle_vs_ref.py, run wherever oracle/_ref/libmlvfs_ref.so exists) and against the
 * fixtures under tests/golden/ that were generated from that build.
 *
 * Each function cites the reference file:line it restates.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stddef.h>
#include <stdint.h>

#define ORC_EV_RES     32768          /* mlvfs.h:87 EV_RESOLUTION */
#define ORC_MAX_BLACK  16384          /* mlvfs.h:88 */

/* ---- LUTs (main.c:128-196) ---- */
/* raw2ev[v], valid for 0 <= v < 16384 + black; returns pointer into a static table, NULL if black > 16384 */
const int    *orc_raw2ev(int black);
const double *orc_raw2evf(int black);
/* ev2raw[e], valid for -10*EV <= e < 14*EV */
const int    *orc_ev2raw(void);

/* ---- unpack (dng.c:813-872) ---- */
size_t orc_unpack(const uint16_t *packed, uint8_t *out, long offset, size_t max_size, int bpp);

/* ---- chroma smoothing (cs.c:49-84, chroma_smooth.c:22-71) ---- */
void orc_chroma_smooth_u16(uint16_t *img, int w, int h, int black, int method);
/* 20-bit twin used by dual ISO (hdr.c:1488-1522): in -> out, caller-provided LUTs, black = 0 */
void orc_chroma_smooth_u32(const uint32_t *inp, uint32_t *out, int w, int h, int method,
                           const int *raw2ev, const int *ev2raw);

/* ---- bad / focus pixels (cs.c:87-331, 440-503) ---- */
typedef struct { int x, y; } orc_pixel;
/* detection pass of fix_bad_pixels (cs.c:257-306); returns count, writes up to cap entries */
size_t orc_badpix_detect(const uint16_t *img, int w, int h, int black, int aggressive,
                         int crop_x, int crop_y, orc_pixel *list, size_t cap);
/* application pass (cs.c:314-330) */
void orc_badpix_apply(uint16_t *img, int w, int h, int black, const orc_pixel *list, size_t n,
                      int crop_x, int crop_y, int dual_iso);
/* fix_focus_pixels body (cs.c:462-501) for an already-loaded map */
void orc_focuspix_apply(uint16_t *img, int w, int h, int black, const orc_pixel *map, size_t n,
                        int crop_x, int crop_y, int dual_iso);

/* ---- vertical stripes (stripes.c:102-266) ---- */
typedef struct { uint32_t r[34]; int f, b; int primed; } orc_rand_t;   /* glibc TYPE_3 rand() */
void orc_rand_seed(orc_rand_t *st, unsigned seed);
int  orc_rand_next(orc_rand_t *st);
/* returns correction_needed; coef[8] written like stripes_compute_correction (coef of skipped groups keep input) */
int  orc_stripes_compute(const uint16_t *img, int w, int h, int black, int white, int frame_size,
                         orc_rand_t *rng, int coef[8]);
void orc_stripes_apply(uint16_t *img, size_t n, int w, int black, int white, int needed, const int coef[8]);

/* ---- LJ92 (lj92.c:82-709, main.c:617-681) ---- */
int  orc_lj92_decode(const uint8_t *data, int len, uint16_t *out, int cap, int *w, int *h, int *bits);
void orc_lj92_untile(const uint16_t *src, uint16_t *dst, int w, int h);
long orc_lj92_encode(const uint16_t *img, int w, int h, int depth, uint8_t *out, size_t cap);

/* ---- pattern noise (patternnoise.c:47-380) ---- */
void orc_fix_pattern_noise(int16_t *raw, int w, int h, int white);

/* ---- dual ISO (hdr.c:250-1957) ---- */
typedef struct {                 /* the reference's function-static LUT state */
    int *raw2ev, *ev2raw_0;      /* 20-bit tables (hdr.c:839-874) */
    int lut_black, lut_white;
    double *fullres_curve;       /* hdr.c:890-913 */
    int curve_black;
    int amaze_fresh_tiles;       /* test knob, see orc_amaze.c */
} orc_diso_state;
typedef struct {                 /* intermediate results, for stage-level parity checks */
    int rggb, is_bright[4], white_dark, white_bright, white_darkened;
    double a, b, corr_ev, overlap;
} orc_diso_info;
void orc_diso_state_init(orc_diso_state *S);
void orc_diso_state_free(orc_diso_state *S);
int  orc_hdr_check(const uint16_t *img, int w, int h, int black, int white);
int  orc_hdr_interpolate(uint16_t *image, int w, int h, int black14, int interp_method, int use_fullres,
                         int use_alias_map, int cs_method, orc_diso_state *S, orc_diso_info *info);

/* ---- AMaZE demosaic, SSE2 variant (amaze_demosaic_RT.c:113-1487) ---- */
void orc_amaze_demosaic(const float *raw, float *red, float *green, float *blue, int stride, int width, int height,
                        int fresh_tiles);

/* ---- dual-ISO preview (hdr.c:40-227) and deflicker (main.c:895-906, histogram.c) ---- */
int  orc_hdr_preview(uint16_t *img, int width, int height, int black_level, int white_level, size_t max_size,
                     const orc_pixel *focus, size_t nfocus, int crop_x, int crop_y);
void orc_deflicker(const uint16_t *img, size_t bytes, int bpp, int black_level, int target, int bias[2]);

/* ---- whole single-ISO chain in process_frame order (main.c:942-997) ---- */
typedef struct {
    int chroma_smooth;      /* 0,2,3,5 */
    int fix_bad_pixels;     /* 0,1,2 */
    int fix_stripes;        /* 0,1 */
    int fix_pattern_noise;  /* 0,1 */
} orc_opts;

#endif
