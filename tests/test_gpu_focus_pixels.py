"""GPU parity: fix_focus_pixels (cs.c:440-503) -- the map loader (cs.c:355-402), the interior interpolators, the
border rules (cs.c:479-500: vertical / horizontal / copy variants) and entries that fall outside [0, w) but still act
on the wrapped linear index.  Bit-exact against the oracle (pinned to the compiled reference) and against the golden
fixture made from one of the reference's real maps."""
import ctypes as C
import os

import numpy as np
import pytest

import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F, synth

from test_focus_golden import CAM, CASES, G, RAW, golden_frame

pytestmark = pytest.mark.gpu


def write_fpm(xy, cam, rw, rh):
    with open("%x_%dx%d.fpm" % (cam, rw, rh), "w") as f:
        for x, y in xy:
            f.write(f"{int(x)} \t {int(y)}\n")


@pytest.fixture()
def in_tmp(tmp_path):
    cwd = os.getcwd()
    os.chdir(tmp_path)
    yield tmp_path
    os.chdir(cwd)


def synthetic_map(w, h, crop, rng):
    pts = [(int(x) + crop[0], int(y) + crop[1]) for x, y in zip(rng.integers(-6, w + 6, 900), rng.integers(-3, h + 3, 900))]
    pts += [(x + crop[0], 50 + crop[1]) for x in range(100, 130, 2)]                      # a dependent horizontal run
    pts += [(77 + crop[0], y + crop[1]) for y in range(60, 90, 2)]                        # a dependent vertical run
    pts += [(0 + crop[0], 5 + crop[1]), (w - 1 + crop[0], 7 + crop[1]), (3 + crop[0], 1 + crop[1]), (200 + crop[0], h - 1 + crop[1])]
    pts += [(x + crop[0], y + crop[1]) for y in (0, 1, 2, 3, h - 3, h - 2, h - 1) for x in (-3, -2, -1, w, w + 1, w + 2)]   # wrapped
    pts += [(x + crop[0], 2 + crop[1]) for x in range(w - 6, w + 5)]                      # a chain across the row end
    pts += [(30 + crop[0], 40 + crop[1])] * 2                                             # duplicate entry
    return pts


@pytest.mark.parametrize("dual", [0, 1])
def test_focus_pixels_synthetic_map_matches_oracle(fresh_ctx, oracle, in_tmp, dual):
    w, h = 400, 200
    cam = 0x80000331
    hdr = F.make_frame_headers(w, h, camera_model=cam, raw_width=1808, raw_height=727, pan_x=16, pan_y=10)
    crop = ((16 + 7) & ~7, 10 & ~1)
    pts = synthetic_map(w, h, crop, np.random.default_rng(9))
    write_fpm(pts, cam, 1808, 727)
    for seed in (4, 5):                                                                   # second frame: cached map + schedule
        img = synth.make_frame(w, h, seed)
        want = oracle.focuspix_apply(img, 2048, np.array(pts, np.int32), crop=crop, dual_iso=dual)
        got = M.fix_focus_pixels(hdr, img.copy(), dual)
        assert (want != img).sum() > 300
        assert np.array_equal(got, want), int(np.count_nonzero(got != want))


@pytest.mark.parametrize("k", [0, 1])
@pytest.mark.parametrize("dual", [0, 1])
def test_focus_pixels_real_map_matches_reference_golden(fresh_ctx, in_tmp, k, dual):
    w, h, px, py, _ = CASES[k]
    write_fpm(G["fpm_xy"], CAM, *RAW)
    hdr = F.make_frame_headers(w, h, camera_model=CAM, raw_width=RAW[0], raw_height=RAW[1], pan_x=px, pan_y=py)
    img, want, _ = golden_frame(k, dual)
    got = M.fix_focus_pixels(hdr, img.copy(), dual)
    assert np.array_equal(got, want), int(np.count_nonzero(got != want))


def test_focus_pixels_in_the_frame_pipeline(fresh_ctx, oracle, in_tmp):
    """process_frame order (main.c:966-973): focus pixels, then bad pixels, then chroma smoothing; the panned
    geometry changes between frames (a schedule per crop offset)."""
    w, h = 640, 360
    cam = 0x80000346
    rng = np.random.default_rng(3)
    base = [(int(x), int(y)) for x, y in zip(rng.integers(0, 1808, 4000), rng.integers(0, 727, 4000))]
    write_fpm(base, cam, 1808, 727)
    o = M.Options(chroma_smooth=2, fix_bad_pixels=0)
    for i, (px, py) in enumerate([(100, 20), (100, 20), (333, 51)]):
        hdr = F.make_frame_headers(w, h, camera_model=cam, raw_width=1808, raw_height=727, pan_x=px, pan_y=py,
                                   file_guid=0xF0C05)
        crop = ((px + 7) & ~7, py & ~1)
        img = synth.make_frame(w, h, i)
        want = oracle.chroma_smooth(oracle.focuspix_apply(img, 2048, np.array(base, np.int32), crop=crop), 2048, 2)
        got, res = fresh_ctx.process_frame(hdr, synth.pack_bits(img), o, "focus.MLV")
        assert res.status == 0
        assert np.array_equal(got, want), (i, int(np.count_nonzero(got != want)))


def test_focus_pixels_dual_iso_conversion(fresh_ctx, oracle, in_tmp):
    """cr2hdr20_convert_data repairs focus pixels with the horizontal interpolator first (hdr.c:1944-1948)."""
    w, h = 640, 360
    cam = 0x80000301
    rng = np.random.default_rng(8)
    pts = [(int(x), int(y)) for x, y in zip(rng.integers(-2, w + 2, 1500), rng.integers(0, h, 1500))]
    write_fpm(pts, cam, 1808, 727)
    hdr = F.make_frame_headers(w, h, camera_model=cam, raw_width=1808, raw_height=727, file_guid=0xF0C06)
    img = synth.make_frame(w, h, 2, dual_iso=True)
    rc, want, _ = oracle.cr2hdr20(img, 2048, 15000, interp_method=1, chroma_smooth_method=0,
                                  focus_map=np.array(pts, np.int32))
    assert rc == 1
    got = img.copy()
    L = M.lib()
    L.cr2hdr20_convert_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    assert L.cr2hdr20_convert_data(C.byref(hdr), got.ctypes.data_as(C.c_void_p), 1, 1, 1, 0, 0) == 1
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1, int(d.max())
