/*
 * dng_header.c -- the CinemaDNG header of a virtual frame: dng_get_header_data / dng_get_header_size, the step
 * right after the pixel path (reference dng.c:612-803).  Host C: a few hundred bytes of
 * TIFF directory per frame, written after the GPU stages because dual ISO changes black / white levels
 * (main.c:961-965) and deflicker sets the baseline exposure (main.c:895-906).
 *
 * The output is byte-identical to the reference's (tests/test_dng_header.py compares the whole 64 KiB block with
 * the compiled reference's for several cameras / white-balance modes / crop cases), which fixes: the order in which the
 * out-of-line values are laid out behind the two directories, the in-directory packing of strings of up to 4
 * bytes, the float / double mix of the white-balance computation (a dcraw-style matrix pseudo-inverse, dng.c:
 * 287-446) and the time-code / date formatting.  Camera calibration data: dng_camera_tables.h.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dng_camera_tables.h"
#include "frame_builder.h"

#define HEADER_BYTES MLVB_DNG_HEADER_SIZE
#define N_IFD0 41
#define N_EXIF 11

enum { T_BYTE = 1, T_ASCII = 2, T_SHORT = 3, T_LONG = 4, T_RATIONAL = 5, T_UNDEFINED = 7, T_SRATIONAL = 10 };

/* ---- TIFF directory writer: entries in call order, out-of-line data appended behind both directories ---- */
struct tiff_out {
    uint8_t *buf;
    uint32_t data;        /* next free byte of the data area */
    uint32_t entry;       /* where the next 12-byte entry goes */
};

static void put_entry(struct tiff_out *t, uint16_t tag, uint16_t type, uint32_t count, uint32_t value)
{
    uint8_t *e = t->buf + t->entry;
    memcpy(e, &tag, 2);
    memcpy(e + 2, &type, 2);
    memcpy(e + 4, &count, 4);
    memcpy(e + 8, &value, 4);
    t->entry += 12;
}

static uint32_t put_words(struct tiff_out *t, const int32_t *v, int n)
{
    const uint32_t at = t->data;
    memcpy(t->buf + at, v, (size_t)n * 4);
    t->data += (uint32_t)n * 4;
    return at;
}

static void entry_ints(struct tiff_out *t, uint16_t tag, uint16_t type, uint32_t count, const int32_t *v, int nwords)
{
    put_entry(t, tag, type, count, put_words(t, v, nwords));
}

static void entry_rational(struct tiff_out *t, uint16_t tag, uint16_t type, int32_t num, int32_t den)
{
    const int32_t v[2] = {num, den};
    put_entry(t, tag, type, 1, put_words(t, v, 2));
}

/* strings of up to 4 bytes (with the NUL) live in the entry itself; longer ones go to the data area, 2-aligned */
static void entry_string(struct tiff_out *t, uint16_t tag, const char *s)
{
    const size_t len = strlen(s) + 1;
    uint32_t value = 0;
    if (len <= 4) memcpy(&value, s, len);
    else {
        value = t->data;
        memcpy(t->buf + value, s, len);
        t->data += (uint32_t)len;
        if (t->data % 2) t->data++;
    }
    put_entry(t, tag, T_ASCII, (uint32_t)len, value);
}

/* ---- white balance: kelvin -> as-shot neutral through the camera's ColorMatrix2 (dng.c:273-446) ---- */

static const double k_xyz_to_rgb[3][3] = {{3.24071, -0.969258, 0.0556352}, {-1.53726, 1.87599, -0.203996}, {-0.498571, 0.0415557, 1.05707}};
static const double k_rgb_to_xyz[3][3] = {{0.412453, 0.357580, 0.180423}, {0.212671, 0.715160, 0.072169}, {0.019334, 0.119193, 0.950227}};

/* daylight-locus chromaticity for a colour temperature, then to linear RGB, normalised to max 1 */
static void daylight_rgb(double T, double rgb[3])
{
    double x;
    if (T <= 4000) x = 0.27475e9 / (T * T * T) - 0.98598e6 / (T * T) + 1.17444e3 / T + 0.145986;
    else if (T <= 7000) x = -4.6070e9 / (T * T * T) + 2.9678e6 / (T * T) + 0.09911e3 / T + 0.244063;
    else x = -2.0064e9 / (T * T * T) + 1.9018e6 / (T * T) + 0.24748e3 / T + 0.237040;
    const double y = -3 * x * x + 2.87 * x - 0.275;
    const double X = x / y, Y = 1, Z = (1 - x - y) / y;
    double top = 0;
    for (int c = 0; c < 3; c++) {
        rgb[c] = X * k_xyz_to_rgb[0][c] + Y * k_xyz_to_rgb[1][c] + Z * k_xyz_to_rgb[2][c];
        if (rgb[c] > top) top = rgb[c];
    }
    for (int c = 0; c < 3; c++) rgb[c] = rgb[c] / top;
}

/* Moore-Penrose pseudo-inverse of a (rows x 3) matrix by Gauss-Jordan on the normal equations */
static void pinv_nx3(double (*in)[3], double (*out)[3], int rows)
{
    double w[3][6];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 6; j++) w[i][j] = j == i + 3;
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < rows; k++) w[i][j] += in[k][i] * in[k][j];
    }
    for (int i = 0; i < 3; i++) {
        double piv = w[i][i];
        for (int j = 0; j < 6; j++) w[i][j] /= piv;
        for (int k = 0; k < 3; k++) {
            if (k == i) continue;
            piv = w[k][i];
            for (int j = 0; j < 6; j++) w[k][j] -= w[i][j] * piv;
        }
    }
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < 3; j++) {
            out[i][j] = 0;
            for (int k = 0; k < 3; k++) out[i][j] += w[j][k + 3] * in[i][k];
        }
}

static void kelvin_to_multipliers(double kelvin, double green, double mul[3], const struct dng_camera_matrices *cam)
{
    double cam_xyz[3][3], cam_rgb[3][3], inv[3][3], rgb_cam_t[3][3], back[3][3], wb[3];
    float pre_mul[3], rgb_cam[3][3];                     /* single precision on purpose: the reference keeps these as float */
    for (int i = 0; i < 9; i++) cam_xyz[i / 3][i % 3] = (double)cam->m[1][i] / (double)10000;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            cam_rgb[i][j] = 0;
            for (int k = 0; k < 3; k++) cam_rgb[i][j] += cam_xyz[i][k] * k_rgb_to_xyz[k][j];
        }
    for (int i = 0; i < 3; i++) {                        /* rows sum to 1 */
        double sum = 0;
        for (int j = 0; j < 3; j++) sum += cam_rgb[i][j];
        for (int j = 0; j < 3; j++) cam_rgb[i][j] /= sum;
        pre_mul[i] = (float)(1 / sum);
    }
    pinv_nx3(cam_rgb, inv, 3);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) rgb_cam[i][j] = (float)inv[j][i];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) rgb_cam_t[i][j] = rgb_cam[j][i];
    pinv_nx3(rgb_cam_t, back, 3);
    daylight_rgb(kelvin, wb);
    wb[1] = wb[1] / green;
    for (int c = 0; c < 3; c++) {
        double inv_mul = 0;
        for (int cc = 0; cc < 3; cc++) inv_mul += (1 / pre_mul[c]) * back[c][cc] * wb[cc];     /* 1 / float: a float quotient */
        mul[c] = 1 / inv_mul;
    }
    mul[0] /= mul[1];
    mul[2] /= mul[1];
    mul[1] = 1;
}

static void as_shot_neutral(const mlv_wbal_hdr_t *wb, int32_t out[6], const struct dng_camera_matrices *cam)
{
    enum { WB_AUTO = 0, WB_SUNNY = 1, WB_CLOUDY = 2, WB_TUNGSTEN = 3, WB_FLUORESCENT = 4, WB_FLASH = 5, WB_CUSTOM = 6, WB_SHADE = 8, WB_KELVIN = 9 };
    if (wb->wb_mode == WB_CUSTOM) {
        out[0] = (int32_t)wb->wbgain_r; out[1] = (int32_t)wb->wbgain_g;
        out[2] = (int32_t)wb->wbgain_g; out[3] = (int32_t)wb->wbgain_g;
        out[4] = (int32_t)wb->wbgain_b; out[5] = (int32_t)wb->wbgain_g;
        return;
    }
    double kelvin = 5500;
    switch (wb->wb_mode) {
    case WB_AUTO: case WB_KELVIN: kelvin = wb->kelvin; break;
    case WB_SHADE: kelvin = 7000; break;
    case WB_CLOUDY: kelvin = 6000; break;
    case WB_TUNGSTEN: kelvin = 3200; break;
    case WB_FLUORESCENT: kelvin = 4000; break;
    default: break;                                      /* sunny, flash, unknown modes: 5500 K */
    }
    double mul[3];
    kelvin_to_multipliers(kelvin, 1.0, mul, cam);
    for (int c = 0; c < 3; c++) { out[2 * c] = 1000000; out[2 * c + 1] = (int32_t)(mul[c] * 1000000); }
}

/* ---- small formatters ---- */

static uint8_t bcd(int v) { return (uint8_t)(((v / 10) << 4) | (v % 10)); }

/* SMPTE time code of frame `frame` at the frame rate rounded up to an integer rate (dng.c:532-571) */
static uint32_t put_timecode(struct tiff_out *t, double fps, int frame)
{
    const uint32_t at = t->data;
    memset(t->buf + at, 0, 8);
    const double seconds_f = fps == 0 ? 0 : frame / (fps > 1 ? round(fps) : fps);
    const int hours = (int)floor(seconds_f / 3600), minutes = ((int)floor(seconds_f / 60)) % 60, seconds = ((int)floor(seconds_f)) % 60;
    const int frames = fps > 1 ? (frame % ((int)round(fps))) : 0;
    t->buf[at] = bcd(frames) & 0x3F;
    t->buf[at + 1] = bcd(seconds) & 0x7F;
    t->buf[at + 2] = bcd(minutes) & 0x7F;
    t->buf[at + 3] = bcd(hours) & 0x3F;
    t->data += 8;
    return at;
}

/* recording start (RTCI) + the frame's offset into the clip; days are not carried into the month (dng.c:583-598) */
static void frame_datetime(char *out, size_t cap, const struct frame_headers *fh)
{
    const uint32_t s = fh->rtci_hdr.tm_sec + (uint32_t)((fh->vidf_hdr.timestamp - fh->rtci_hdr.timestamp) / 1000000);
    const uint32_t m = fh->rtci_hdr.tm_min + s / 60, h = fh->rtci_hdr.tm_hour + m / 60, d = fh->rtci_hdr.tm_mday + h / 24;
    snprintf(out, cap, "%04d:%02d:%02d %02d:%02d:%02d", 1900 + fh->rtci_hdr.tm_year, fh->rtci_hdr.tm_mon + 1, d, h % 24, m % 60, s % 60);
}

static void bounded_copy(char *dst, size_t cap, const uint8_t *src, size_t n)
{
    if (n >= cap) n = cap - 1;
    memcpy(dst, src, n);
    dst[n] = 0;
}

size_t dng_get_header_size(void) { return HEADER_BYTES; }

size_t dng_get_header_data(struct frame_headers *fh, uint8_t *output_buffer, off_t offset, size_t max_size, double fps_override,
                           char *mlv_basename)
{
    uint8_t *hdr = calloc(1, HEADER_BYTES);
    if (!hdr) return 0;
    struct raw_info *ri = &fh->rawi_hdr.raw_info;

    char model[33], make[33], serial[33], lens[33], datetime[64];
    bounded_copy(model, sizeof(model), fh->idnt_hdr.cameraName, strnlen((const char *)fh->idnt_hdr.cameraName, 32));
    snprintf(make, sizeof(make), "%s", model);                   /* the make is the first word of the camera name */
    char *space = strchr(make, ' ');
    if (space) *space = 0;
    bounded_copy(serial, sizeof(serial), fh->idnt_hdr.cameraSerial, strnlen((const char *)fh->idnt_hdr.cameraSerial, 32));
    bounded_copy(lens, sizeof(lens), fh->lens_hdr.lensName, strnlen((const char *)fh->lens_hdr.lensName, 32));
    frame_datetime(datetime, sizeof(datetime), fh);

    /* calibration rows: the camera's own, else the first row of each table */
    const struct dng_camera_matrices *cam = &dng_camera_matrices[0];
    for (size_t i = 0; i < sizeof(dng_camera_matrices) / sizeof(dng_camera_matrices[0]); i++)
        if (!strcmp(dng_camera_matrices[i].camera, model)) { cam = &dng_camera_matrices[i]; break; }
    const struct dng_camera_focal *foc = &dng_camera_focal[0];
    for (size_t i = 0; i < sizeof(dng_camera_focal) / sizeof(dng_camera_focal[0]); i++)
        if (!strcmp(dng_camera_focal[i].camera, model)) { foc = &dng_camera_focal[i]; break; }
    int32_t focal_x[2] = {foc->x[0], foc->x[1]}, focal_y[2] = {foc->y[0], foc->y[1]};

    /* pixel aspect: a raw buffer wider than 2:1 and at most 720 rows is 5x3 line-skipped footage; other buffers
       narrower than 2000 px come from 3x3 binning modes (dng.c:654-676) */
    int32_t scale[4] = {1, 1, 1, 1};
    const double raw_w = ri->active_area.x2 - ri->active_area.x1, raw_h = ri->active_area.y2 - ri->active_area.y1;
    if (raw_w / raw_h > 2.0 && raw_h <= 720) {
        scale[2] = 5; scale[3] = 3;
        focal_x[1] *= 3; focal_y[1] *= 5;
    } else if (raw_w < 2000) {
        focal_x[1] *= 3; focal_y[1] *= 3;
    }
    /* the active area describes the camera's raw buffer; a recording that does not contain the optical-black
       borders gets the recorded frame as its active area -- written back into the caller's headers like dng.c:678-687 */
    if (fh->rawi_hdr.xRes < ri->active_area.x2 || fh->rawi_hdr.yRes < ri->active_area.y2) {
        ri->active_area.x1 = 0; ri->active_area.y1 = 0;
        ri->active_area.x2 = fh->rawi_hdr.xRes; ri->active_area.y2 = fh->rawi_hdr.yRes;
    }
    int32_t rate[2] = {(int32_t)fh->file_hdr.sourceFpsNom, (int32_t)fh->file_hdr.sourceFpsDenom};
    if (fps_override > 0) { rate[0] = (int32_t)fps_override * 1000; rate[1] = 1000; }
    const double rate_f = rate[1] == 0 ? 0 : (double)rate[0] / (double)rate[1];
    int32_t baseline[2] = {ri->exposure_bias[0], ri->exposure_bias[1]};
    if (baseline[1] == 0) { baseline[0] = 0; baseline[1] = 1; }
    int32_t neutral[6];
    as_shot_neutral(&fh->wbal_hdr, neutral, cam);
    int32_t mats[4][18];
    for (int m = 0; m < 4; m++)
        for (int i = 0; i < 9; i++) { mats[m][2 * i] = cam->m[m][i]; mats[m][2 * i + 1] = 10000; }

    const uint32_t ifd0_at = 8, exif_at = ifd0_at + 2 + N_IFD0 * 12 + 4, data_at = exif_at + 2 + N_EXIF * 12 + 4;
    const uint16_t tiff_magic[4] = {0x4949, 42, 8, 0};           /* "II", 42, first directory at byte 8 */
    memcpy(hdr, tiff_magic, sizeof(tiff_magic));
    struct tiff_out t = {hdr, data_at, ifd0_at + 2};
    const uint16_t n0 = N_IFD0, n1 = N_EXIF;
    memcpy(hdr + ifd0_at, &n0, 2);

    const uint32_t crop_origin = (uint32_t)((((uint16_t)ri->crop.origin[1]) << 16) | ((uint16_t)ri->crop.origin[0]));
    const uint32_t crop_size = (uint32_t)((((uint16_t)(ri->active_area.y2 - ri->active_area.y1)) << 16) |
                                          ((uint16_t)(ri->active_area.x2 - ri->active_area.x1)));
    put_entry(&t, 254, T_LONG, 1, 0);                                             /* NewSubFileType: main image */
    put_entry(&t, 256, T_LONG, 1, fh->rawi_hdr.xRes);
    put_entry(&t, 257, T_LONG, 1, fh->rawi_hdr.yRes);
    put_entry(&t, 258, T_SHORT, 1, 16);                                           /* BitsPerSample */
    put_entry(&t, 259, T_SHORT, 1, 1);                                            /* uncompressed */
    put_entry(&t, 262, T_SHORT, 1, 32803);                                        /* colour filter array */
    put_entry(&t, 266, T_SHORT, 1, 1);                                            /* FillOrder */
    entry_string(&t, 271, make);
    entry_string(&t, 272, model);
    put_entry(&t, 273, T_LONG, 1, HEADER_BYTES);                                  /* the strip starts right after the header */
    put_entry(&t, 274, T_SHORT, 1, 1);
    put_entry(&t, 277, T_SHORT, 1, 1);
    put_entry(&t, 278, T_SHORT, 1, fh->rawi_hdr.yRes);
    put_entry(&t, 279, T_LONG, 1, (uint32_t)dng_get_image_size(fh));
    put_entry(&t, 284, T_SHORT, 1, 1);
    entry_string(&t, 305, "MLVFS");
    entry_string(&t, 306, datetime);
    put_entry(&t, 33421, T_SHORT, 2, 0x00020002);                                 /* CFA repeat 2x2 */
    put_entry(&t, 33422, T_BYTE, 4, 0x02010100);                                  /* RGGB */
    put_entry(&t, 34665, T_LONG, 1, exif_at);
    put_entry(&t, 50706, T_BYTE, 4, 0x00000401);                                  /* DNG 1.4.0.0 */
    entry_string(&t, 50708, model);
    put_entry(&t, 50714, T_LONG, 1, (uint32_t)ri->black_level);
    put_entry(&t, 50717, T_LONG, 1, (uint32_t)ri->white_level);
    entry_ints(&t, 50718, T_RATIONAL, 2, scale, 4);
    put_entry(&t, 50719, T_SHORT, 2, crop_origin);
    put_entry(&t, 50720, T_SHORT, 2, crop_size);
    entry_ints(&t, 50721, T_SRATIONAL, 9, mats[0], 18);
    entry_ints(&t, 50722, T_SRATIONAL, 9, mats[1], 18);
    entry_ints(&t, 50728, T_RATIONAL, 3, neutral, 6);
    entry_ints(&t, 50730, T_SRATIONAL, 1, baseline, 2);
    entry_string(&t, 50735, serial);
    put_entry(&t, 50778, T_SHORT, 1, 17);                                         /* standard light A */
    put_entry(&t, 50779, T_SHORT, 1, 21);                                         /* D65 */
    entry_ints(&t, 50829, T_LONG, 4, ri->dng_active_area, 4);
    entry_ints(&t, 50964, T_SRATIONAL, 9, mats[2], 18);
    entry_ints(&t, 50965, T_SRATIONAL, 9, mats[3], 18);
    {
        const uint32_t at = put_timecode(&t, rate_f, (int)fh->vidf_hdr.frameNumber);
        put_entry(&t, 51043, T_BYTE, 8, at);                                      /* CinemaDNG TimeCodes */
    }
    entry_ints(&t, 51044, T_SRATIONAL, 1, rate, 2);                               /* CinemaDNG FrameRate */
    entry_string(&t, 51081, mlv_basename ? mlv_basename : "");                    /* CinemaDNG ReelName */
    entry_rational(&t, 51109, T_SRATIONAL, 0, 1);                                 /* BaselineExposureOffset */
    /* next-directory link of IFD0 stays 0 */

    memcpy(hdr + exif_at, &n1, 2);
    t.entry = exif_at + 2;
    entry_rational(&t, 33434, T_RATIONAL, (int32_t)fh->expo_hdr.shutterValue / 1000, 1000);
    entry_rational(&t, 33437, T_RATIONAL, fh->lens_hdr.aperture, 100);
    put_entry(&t, 34855, T_SHORT, 1, fh->expo_hdr.isoValue);
    put_entry(&t, 34864, T_SHORT, 1, 3);                                          /* sensitivity type: ISO speed */
    put_entry(&t, 36864, T_UNDEFINED, 4, 0x30333230);                             /* Exif "0230" */
    entry_rational(&t, 37382, T_RATIONAL, (int32_t)fh->lens_hdr.focalDist, 1);
    entry_rational(&t, 37386, T_RATIONAL, fh->lens_hdr.focalLength, 1);
    entry_ints(&t, 41486, T_RATIONAL, 1, focal_x, 2);
    entry_ints(&t, 41487, T_RATIONAL, 1, focal_y, 2);
    put_entry(&t, 41488, T_SHORT, 1, (uint32_t)foc->unit);
    entry_string(&t, 42036, lens);

    size_t n = 0;
    if (offset >= 0 && (size_t)offset < HEADER_BYTES) {
        n = max_size < HEADER_BYTES - (size_t)offset ? max_size : HEADER_BYTES - (size_t)offset;
        if (n) memcpy(output_buffer, hdr + offset, n);
    }
    free(hdr);
    return n;
}
