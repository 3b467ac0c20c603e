"""The host DNG header writer (mlvfs_b200/host/dng_header.c: dng_get_header_data, reference dng.c:612-803), the whole
64 KiB block byte for byte against the compiled reference: cameras with and without calibration rows, every
white-balance mode, crop / line-skipping geometries, fps override, short strings packed into the directory, the
time code and date arithmetic.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mlvfs_b200 import mlvformat as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "mlvfs_b200", "libmlvfs_b200_host.so")
SIG = [C.c_void_p, C.c_void_p, C.c_long, C.c_size_t, C.c_double, C.c_char_p]


@pytest.fixture(scope="module")
def host():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mlvfs_b200", "host")])
    lib = C.CDLL(HOST_SO)
    lib.dng_get_header_data.restype = C.c_size_t
    lib.dng_get_header_data.argtypes = SIG
    lib.dng_get_header_size.restype = C.c_size_t
    return lib


def headers(camera="Canon EOS 5D Mark III", w=1920, h=1080, raw_w=None, raw_h=None, wb_mode=0, kelvin=5200, frame=7,
            serial=b"1A2B3C4D5E", lens=b"EF24-70mm f/2.8L II USM", black=2048, white=15000, bias=(0, 1), fps=(24000, 1001),
            gains=(2100, 1024, 1600), active=None, crop_origin=(0, 0)):
    fh = F.make_frame_headers(w, h, raw_width=raw_w, raw_height=raw_h, camera_name=camera, black=black, white=white,
                              frame_number=frame)
    ri = fh.rawi_hdr.raw_info
    if active is not None:
        ri.active_area[0], ri.active_area[1], ri.active_area[2], ri.active_area[3] = active
    ri.exposure_bias[0], ri.exposure_bias[1] = bias
    ri.crop[0], ri.crop[1] = crop_origin
    fh.file_hdr.sourceFpsNom, fh.file_hdr.sourceFpsDenom = fps
    C.memmove(fh.idnt_hdr.cameraSerial, serial, len(serial))
    fh.vidf_hdr.timestamp = 3_700_000_000 + frame * 41708
    F._tag(fh.rtci_hdr.blockType, "RTCI")
    fh.rtci_hdr.timestamp = 12345
    for i, v in enumerate((58, 59, 23, 28, 1, 124)):          # sec, min, hour, mday, mon, year: rolls over midnight
        fh.rtci_hdr.tm[i] = v
    F._tag(fh.expo_hdr.blockType, "EXPO")
    fh.expo_hdr.isoValue = 800
    fh.expo_hdr.shutterValue = 20833
    F._tag(fh.lens_hdr.blockType, "LENS")
    fh.lens_hdr.focalLength, fh.lens_hdr.focalDist, fh.lens_hdr.aperture = 50, 1200, 280
    C.memmove(fh.lens_hdr.lensName, lens, len(lens))
    F._tag(fh.wbal_hdr.blockType, "WBAL")
    fh.wbal_hdr.wb_mode, fh.wbal_hdr.kelvin = wb_mode, kelvin
    fh.wbal_hdr.wbgain_r, fh.wbal_hdr.wbgain_g, fh.wbal_hdr.wbgain_b = gains
    return fh


def both(host, ref, fh, fps_override=0.0, base=b"/M19-1234.MLV"):
    ref.dng_get_header_data.restype = C.c_size_t
    ref.dng_get_header_data.argtypes = SIG
    a, b = F.clone_headers(fh), F.clone_headers(fh)
    out_a, out_b = np.full(65536, 0xEE, np.uint8), np.full(65536, 0xEE, np.uint8)
    na = host.dng_get_header_data(C.byref(a), out_a.ctypes.data, 0, 65536, fps_override, base)
    nb = ref.dng_get_header_data(C.byref(b), out_b.ctypes.data, 0, 65536, fps_override, base)
    assert na == nb == 65536
    assert bytes(a) == bytes(b)                                  # the active-area write-back into the caller's headers
    return out_a, out_b


CAMERAS = ["Canon EOS 5D Mark III", "Canon EOS 5D Mark II", "Canon EOS 7D", "Canon EOS 6D", "Canon EOS 70D", "Canon EOS 60D",
           "Canon EOS 50D", "Canon EOS 550D", "Canon EOS 600D", "Canon EOS 650D", "Canon EOS 700D", "Canon EOS 1100D", "Canon EOS M",
           "Canon EOS 500D", "Some Other Camera", "X"]


@pytest.mark.parametrize("camera", CAMERAS)
def test_header_matches_reference_per_camera(host, ref, camera):
    for wb_mode, kelvin in [(0, 5200), (9, 3100), (9, 9300), (1, 0), (8, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0), (7, 0)]:
        got, want = both(host, ref, headers(camera=camera, wb_mode=wb_mode, kelvin=kelvin))
        assert np.array_equal(got, want), (camera, wb_mode, int(np.flatnonzero(got != want)[0]))


@pytest.mark.parametrize("kw", [
    dict(w=1920, h=1080, raw_w=2080, raw_h=1318),                                   # full raw buffer wider than the recording
    dict(w=1728, h=624, raw_w=1808, raw_h=727, active=(28, 72, 727, 1808)),         # 5x3 line skipping (aspect > 2, <= 720 rows)
    dict(w=2560, h=1090, raw_w=3584, raw_h=1320, active=(28, 146, 1320, 3584)),     # crop mode, wide buffer
    dict(w=5760, h=3240, raw_w=5936, raw_h=3950),
    dict(w=640, h=360, active=(0, 0, 360, 640), crop_origin=(4, 6)),
    dict(bias=(-3125, 10000)), dict(bias=(5, 0)), dict(fps=(25000, 1000)), dict(fps=(0, 0)), dict(fps=(500, 1000), frame=3),
    dict(frame=123456), dict(serial=b"", lens=b""), dict(serial=b"abc", lens=b"L"), dict(lens=b"x" * 32, serial=b"9" * 32),
    dict(black=8192, white=60000),
])
def test_header_matches_reference_geometry_and_metadata(host, ref, kw):
    got, want = both(host, ref, headers(**kw))
    assert np.array_equal(got, want), (kw, int(np.flatnonzero(got != want)[0]))


def test_header_fps_override_reel_name_and_ranges(host, ref):
    fh = headers()
    for fps, base in [(23.976, b"/clip.MLV"), (30.0, b"/a/b/c/A001.MLV"), (0.5, b"/x"), (60.9, b"")]:
        got, want = both(host, ref, fh, fps, base)
        assert np.array_equal(got, want), (fps, base)
    whole, _ = both(host, ref, fh)
    out = np.full(5000, 0xEE, np.uint8)
    assert host.dng_get_header_data(C.byref(F.clone_headers(fh)), out.ctypes.data, 300, 4096, 0.0, b"/M19-1234.MLV") == 4096
    assert np.array_equal(out[:4096], whole[300:4396]) and np.all(out[4096:] == 0xEE)
    assert host.dng_get_header_data(C.byref(F.clone_headers(fh)), out.ctypes.data, 65536 - 100, 4096, 0.0, b"/M19-1234.MLV") == 100
    assert host.dng_get_header_size() == 65536
