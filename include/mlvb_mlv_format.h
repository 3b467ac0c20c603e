/*
 * mlvb_mlv_format.h -- binary layout of the Magic Lantern Video (MLV v2.0) blocks that the
 * per-frame raw path consumes, and of `struct frame_headers`, the bundle every reference
 * pixel function takes as its first argument.
 *
 * These are on-disk / ABI layouts, so they must match the reference byte for byte:
 *   block structs ............ reference mlvfs/mlv.h:40-239   (#pragma pack(1))
 *   struct raw_info .......... reference mlvfs/raw.h:166-207  (natural alignment, 64-bit host)
 *   struct frame_headers ..... reference mlvfs/mlvfs.h:51-63  (natural alignment)
 * tests/test_abi_layout.py pins every size/offset below against the compiled reference
 * (oracle/_ref) through ref_offsetof_frame_headers().
 *
 * A reference translation unit can keep including its own mlv.h/raw.h/mlvfs.h and pass its
 * `struct frame_headers *` straight into libmlvfs_b200.so: the type names are the same and the
 * layouts are identical.  Define MLVB_NO_FORMAT_TYPES before including mlvfs_b200.h in such a
 * TU to skip these definitions.
 */
#ifndef MLVB_MLV_FORMAT_H
#define MLVB_MLV_FORMAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* file_hdr.videoClass bits (mlv.h:25-32) */
#define MLVB_VIDEO_CLASS_RAW        0x0001
#define MLVB_VIDEO_CLASS_FLAG_DELTA 0x0040
#define MLVB_VIDEO_CLASS_FLAG_LZMA  0x0080
#define MLVB_VIDEO_CLASS_FLAG_LJ92  0x0100

#ifndef MLVB_NO_FORMAT_TYPES

/* raw.h:166-207 -- 64-bit hosts carry a dummy uint32 where the 32-bit firmware has a pointer */
struct raw_info {
    uint32_t api_version;
    uint32_t do_not_use_this;
    int32_t  height, width, pitch;
    int32_t  frame_size;
    int32_t  bits_per_pixel;
    int32_t  black_level;
    int32_t  white_level;
    union {
        struct { int32_t x, y, width, height; } jpeg;
        struct { int32_t origin[2]; int32_t size[2]; } crop;
    };
    union {
        struct { int32_t y1, x1, y2, x2; } active_area;
        int32_t dng_active_area[4];
    };
    int32_t  exposure_bias[2];
    int32_t  cfa_pattern;
    int32_t  calibration_illuminant1;
    int32_t  color_matrix1[18];
    int32_t  dynamic_range;
};

#pragma pack(push, 1)

typedef struct {                /* generic block prefix, mlv.h:40-44 */
    uint8_t  blockType[4];
    uint32_t blockSize;
    uint64_t timestamp;
} mlv_hdr_t;

typedef struct {                /* "MLVI", mlv.h:46-62 */
    uint8_t  fileMagic[4];
    uint32_t blockSize;
    uint8_t  versionString[8];
    uint64_t fileGuid;
    uint16_t fileNum;
    uint16_t fileCount;
    uint32_t fileFlags;
    uint16_t videoClass;
    uint16_t audioClass;
    uint32_t videoFrameCount;
    uint32_t audioFrameCount;
    uint32_t sourceFpsNom;
    uint32_t sourceFpsDenom;
} mlv_file_hdr_t;

typedef struct {                /* "VIDF", mlv.h:64-75; payload follows after frameSpace bytes */
    uint8_t  blockType[4];
    uint32_t blockSize;
    uint64_t timestamp;
    uint32_t frameNumber;
    uint16_t cropPosX;
    uint16_t cropPosY;
    uint16_t panPosX;
    uint16_t panPosY;
    uint32_t frameSpace;
} mlv_vidf_hdr_t;

typedef struct {                /* "RAWI", mlv.h:86-93 */
    uint8_t  blockType[4];
    uint32_t blockSize;
    uint64_t timestamp;
    uint16_t xRes;
    uint16_t yRes;
    struct raw_info raw_info;
} mlv_rawi_hdr_t;

typedef struct {                /* "EXPO", mlv.h:107-117 */
    uint8_t  blockType[4];
    uint32_t blockSize;
    uint64_t timestamp;
    uint32_t isoMode;
    uint32_t isoValue;
    uint32_t isoAnalog;
    uint32_t digitalGain;
    uint64_t shutterValue;
} mlv_expo_hdr_t;

typedef struct {                /* "LENS", mlv.h:119-132 */
    uint8_t  blockType[4];
    uint32_t blockSize;
    uint64_t timestamp;
    uint16_t focalLength;
    uint16_t focalDist;
    uint16_t aperture;
    uint8_t  stabilizerMode;
    uint8_t  autofocusMode;
    uint32_t flags;
    uint32_t lensID;
    uint8_t  lensName[32];
    uint8_t  lensSerial[32];
} mlv_lens_hdr_t;

typedef struct {                /* "RTCI", mlv.h:134-149 */
    uint8_t  blockType[4];
    uint32_t blockSize;
    uint64_t timestamp;
    uint16_t tm_sec, tm_min, tm_hour, tm_mday, tm_mon, tm_year, tm_wday, tm_yday, tm_isdst, tm_gmtoff;
    uint8_t  tm_zone[8];
} mlv_rtci_hdr_t;

typedef struct {                /* "IDNT", mlv.h:151-158 */
    uint8_t  blockType[4];
    uint32_t blockSize;
    uint64_t timestamp;
    uint8_t  cameraName[32];
    uint32_t cameraModel;
    uint8_t  cameraSerial[32];
} mlv_idnt_hdr_t;

typedef struct {                /* "WBAL", mlv.h:215-227 */
    uint8_t  blockType[4];
    uint32_t blockSize;
    uint64_t timestamp;
    uint32_t wb_mode;
    uint32_t kelvin;
    uint32_t wbgain_r, wbgain_g, wbgain_b;
    uint32_t wbs_gm, wbs_ba;
} mlv_wbal_hdr_t;

typedef struct {                /* index entry, mlv.h:160-165 */
    uint16_t fileNumber;
    uint8_t  empty;
    uint8_t  frameType;         /* 1 = VIDF, 2 = AUDF, 0 = anything else */
    uint64_t frameOffset;
} mlv_xref_t;

typedef struct {                /* "XREF", mlv.h:167-175; entryCount mlv_xref_t follow */
    uint8_t  blockType[4];
    uint32_t blockSize;
    uint64_t timestamp;
    uint32_t frameType;
    uint32_t entryCount;
} mlv_xref_hdr_t;

#pragma pack(pop)

/* mlvfs.h:51-63 -- everything a frame's DNG needs, gathered by the host's header walk */
struct frame_headers {
    uint32_t       fileNumber;
    uint64_t       position;     /* file offset of the VIDF block */
    mlv_vidf_hdr_t vidf_hdr;
    mlv_file_hdr_t file_hdr;
    mlv_rtci_hdr_t rtci_hdr;
    mlv_idnt_hdr_t idnt_hdr;
    mlv_rawi_hdr_t rawi_hdr;
    mlv_expo_hdr_t expo_hdr;
    mlv_lens_hdr_t lens_hdr;
    mlv_wbal_hdr_t wbal_hdr;
};

#endif /* MLVB_NO_FORMAT_TYPES */

#ifdef __cplusplus
}
#endif
#endif
