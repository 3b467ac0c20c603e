"""ctypes mirror of include/mlvb_mlv_format.h (MLV v2.0 block layouts + ``struct frame_headers``).

Reference layouts: mlvfs/mlv.h:40-239 (packed blocks), mlvfs/raw.h:166-207 (``struct raw_info``),
mlvfs/mlvfs.h:51-63 (``struct frame_headers``).  tests/test_abi_layout.py pins these against the
compiled reference.
"""
import ctypes as C

VIDEO_CLASS_RAW = 0x0001
VIDEO_CLASS_FLAG_LZMA = 0x0080
VIDEO_CLASS_FLAG_LJ92 = 0x0100


class RawInfo(C.Structure):
    _fields_ = [
        ("api_version", C.c_uint32),
        ("do_not_use_this", C.c_uint32),
        ("height", C.c_int32),
        ("width", C.c_int32),
        ("pitch", C.c_int32),
        ("frame_size", C.c_int32),
        ("bits_per_pixel", C.c_int32),
        ("black_level", C.c_int32),
        ("white_level", C.c_int32),
        ("crop", C.c_int32 * 4),
        ("active_area", C.c_int32 * 4),  # y1, x1, y2, x2
        ("exposure_bias", C.c_int32 * 2),
        ("cfa_pattern", C.c_int32),
        ("calibration_illuminant1", C.c_int32),
        ("color_matrix1", C.c_int32 * 18),
        ("dynamic_range", C.c_int32),
    ]


class _Packed(C.Structure):
    _pack_ = 1


class MlvHdr(_Packed):
    _fields_ = [("blockType", C.c_uint8 * 4), ("blockSize", C.c_uint32), ("timestamp", C.c_uint64)]


class FileHdr(_Packed):
    _fields_ = [
        ("fileMagic", C.c_uint8 * 4),
        ("blockSize", C.c_uint32),
        ("versionString", C.c_uint8 * 8),
        ("fileGuid", C.c_uint64),
        ("fileNum", C.c_uint16),
        ("fileCount", C.c_uint16),
        ("fileFlags", C.c_uint32),
        ("videoClass", C.c_uint16),
        ("audioClass", C.c_uint16),
        ("videoFrameCount", C.c_uint32),
        ("audioFrameCount", C.c_uint32),
        ("sourceFpsNom", C.c_uint32),
        ("sourceFpsDenom", C.c_uint32),
    ]


class VidfHdr(_Packed):
    _fields_ = [
        ("blockType", C.c_uint8 * 4),
        ("blockSize", C.c_uint32),
        ("timestamp", C.c_uint64),
        ("frameNumber", C.c_uint32),
        ("cropPosX", C.c_uint16),
        ("cropPosY", C.c_uint16),
        ("panPosX", C.c_uint16),
        ("panPosY", C.c_uint16),
        ("frameSpace", C.c_uint32),
    ]


class RawiHdr(_Packed):
    _fields_ = [
        ("blockType", C.c_uint8 * 4),
        ("blockSize", C.c_uint32),
        ("timestamp", C.c_uint64),
        ("xRes", C.c_uint16),
        ("yRes", C.c_uint16),
        ("raw_info", RawInfo),
    ]


class ExpoHdr(_Packed):
    _fields_ = [
        ("blockType", C.c_uint8 * 4),
        ("blockSize", C.c_uint32),
        ("timestamp", C.c_uint64),
        ("isoMode", C.c_uint32),
        ("isoValue", C.c_uint32),
        ("isoAnalog", C.c_uint32),
        ("digitalGain", C.c_uint32),
        ("shutterValue", C.c_uint64),
    ]


class LensHdr(_Packed):
    _fields_ = [
        ("blockType", C.c_uint8 * 4),
        ("blockSize", C.c_uint32),
        ("timestamp", C.c_uint64),
        ("focalLength", C.c_uint16),
        ("focalDist", C.c_uint16),
        ("aperture", C.c_uint16),
        ("stabilizerMode", C.c_uint8),
        ("autofocusMode", C.c_uint8),
        ("flags", C.c_uint32),
        ("lensID", C.c_uint32),
        ("lensName", C.c_uint8 * 32),
        ("lensSerial", C.c_uint8 * 32),
    ]


class RtciHdr(_Packed):
    _fields_ = [
        ("blockType", C.c_uint8 * 4),
        ("blockSize", C.c_uint32),
        ("timestamp", C.c_uint64),
        ("tm", C.c_uint16 * 10),
        ("tm_zone", C.c_uint8 * 8),
    ]


class IdntHdr(_Packed):
    _fields_ = [
        ("blockType", C.c_uint8 * 4),
        ("blockSize", C.c_uint32),
        ("timestamp", C.c_uint64),
        ("cameraName", C.c_uint8 * 32),
        ("cameraModel", C.c_uint32),
        ("cameraSerial", C.c_uint8 * 32),
    ]


class WbalHdr(_Packed):
    _fields_ = [
        ("blockType", C.c_uint8 * 4),
        ("blockSize", C.c_uint32),
        ("timestamp", C.c_uint64),
        ("wb_mode", C.c_uint32),
        ("kelvin", C.c_uint32),
        ("wbgain_r", C.c_uint32),
        ("wbgain_g", C.c_uint32),
        ("wbgain_b", C.c_uint32),
        ("wbs_gm", C.c_uint32),
        ("wbs_ba", C.c_uint32),
    ]


class FrameHeaders(C.Structure):
    """``struct frame_headers`` (mlvfs/mlvfs.h:51-63)."""

    _fields_ = [
        ("fileNumber", C.c_uint32),
        ("position", C.c_uint64),
        ("vidf_hdr", VidfHdr),
        ("file_hdr", FileHdr),
        ("rtci_hdr", RtciHdr),
        ("idnt_hdr", IdntHdr),
        ("rawi_hdr", RawiHdr),
        ("expo_hdr", ExpoHdr),
        ("lens_hdr", LensHdr),
        ("wbal_hdr", WbalHdr),
    ]


def _tag(dst, s):
    for i, ch in enumerate(s.encode("ascii")):
        dst[i] = ch


def make_frame_headers(width, height, *, bpp=14, black=2048, white=15000, file_guid=0x1234567890ABCDEF,
                       camera_model=0x80000285, raw_width=None, raw_height=None, pan_x=0, pan_y=0,
                       video_class=VIDEO_CLASS_RAW, frame_number=0, camera_name="Canon EOS 5D Mark III"):
    """Headers for a synthetic clip, filled the way SURVEY.md section 8(d) describes."""
    fh = FrameHeaders()
    _tag(fh.file_hdr.fileMagic, "MLVI")
    fh.file_hdr.blockSize = C.sizeof(FileHdr)
    _tag(fh.file_hdr.versionString, "v2.0")
    fh.file_hdr.fileGuid = file_guid
    fh.file_hdr.fileNum = 0
    fh.file_hdr.fileCount = 1
    fh.file_hdr.videoClass = video_class
    fh.file_hdr.sourceFpsNom = 24000
    fh.file_hdr.sourceFpsDenom = 1000

    _tag(fh.rawi_hdr.blockType, "RAWI")
    fh.rawi_hdr.blockSize = C.sizeof(RawiHdr)
    fh.rawi_hdr.xRes = width
    fh.rawi_hdr.yRes = height
    ri = fh.rawi_hdr.raw_info
    ri.api_version = 1
    ri.height = raw_height if raw_height is not None else height
    ri.width = raw_width if raw_width is not None else width
    ri.pitch = ri.width * bpp // 8
    ri.frame_size = ri.height * ri.pitch
    ri.bits_per_pixel = bpp
    ri.black_level = black
    ri.white_level = white
    ri.active_area[0] = 0
    ri.active_area[1] = 0
    ri.active_area[2] = ri.height
    ri.active_area[3] = ri.width
    ri.exposure_bias[0] = 0
    ri.exposure_bias[1] = 1
    ri.cfa_pattern = 0x02010100
    ri.calibration_illuminant1 = 1
    cm = [6722, 10000, -635, 10000, -963, 10000, -4287, 10000, 12460, 10000, 2028, 10000,
          -908, 10000, 2162, 10000, 5668, 10000]
    for i, v in enumerate(cm):
        ri.color_matrix1[i] = v
    ri.dynamic_range = 1100

    _tag(fh.idnt_hdr.blockType, "IDNT")
    fh.idnt_hdr.blockSize = C.sizeof(IdntHdr)
    _tag(fh.idnt_hdr.cameraName, camera_name)
    fh.idnt_hdr.cameraModel = camera_model

    _tag(fh.vidf_hdr.blockType, "VIDF")
    fh.vidf_hdr.frameNumber = frame_number
    fh.vidf_hdr.panPosX = pan_x
    fh.vidf_hdr.panPosY = pan_y
    fh.vidf_hdr.cropPosX = pan_x & ~7
    fh.vidf_hdr.cropPosY = pan_y & ~1
    fh.vidf_hdr.frameSpace = 0
    return fh


def clone_headers(fh):
    out = FrameHeaders()
    C.memmove(C.byref(out), C.byref(fh), C.sizeof(FrameHeaders))
    return out
