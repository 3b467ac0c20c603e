"""The persistent wide fused kernel (mlvfs_b200/csrc/fused_wide.cuh: 14-bit unpack + 3x3 median chroma smoothing +
stripe gains in one pass) executed on the host by tests/emu/wide_emu.cpp -- the kernel source compiled unchanged, one
std::thread per lane, a barrier for __syncwarp, an exchange array for the shuffles -- against the oracle's single-ISO
chain, bit for bit.  This pins, without a GPU: the work split (per-strip runs cut at frame ends, and the older equal
segments), the row staging and the byte-permute extraction, the sorted-column medians with the shared pair, the
octave lookup of ev2raw, the stripe gain kept in the high half and the PRMT packing, image borders and strip seams.
Bad-pixel patches: the item lists come from the library's host helper (wide_build_items), the repaired values -- computed by
a separate device kernel in the library -- from the oracle.  tests/test_gpu_single_iso.py -k wide then confirms the
same source on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mlvfs_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV = 32768


# the shipped defaults, and the plain forms they replaced (every -DFW_NO_* switch of fused_wide.cuh)
BUILDS = {"default": [], "plain": ["-DFW_NO_MIDPAIR", "-DFW_NO_PACK_PRMT", "-DFW_NO_EXTRACT_SHL", "-DFW_NO_GAIN_X", "-DFW_NO_CLAMP_X2", "-DFW_NO_ADD_FMA"],
          "r01n": ["-DFW_NO_ADD_FMA", "-DFW_NO_CLAMP_X2"],         # round 1's build: per-sample clamp, plain two-input adds
          "cols8": ["-DFW_COLS_CFG=8"]}         # 8 quad columns per lane (480-pixel strips): built, measured slower, kept honest


@pytest.fixture(scope="module", params=sorted(BUILDS))
def emu(request, tmp_path_factory, oracle):
    so = str(tmp_path_factory.mktemp("emu") / f"libwide_emu_{request.param}.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-shared", "-fPIC", "-pthread", "-DMLVB_HOST_EMU", "-DFW_WARPS_CFG=1"] +
                          BUILDS[request.param] + ["-o", so, os.path.join(ROOT, "tests", "emu", "wide_emu.cpp")])
    lib = C.CDLL(so)
    lib.wide_emu_run.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    olib = oracle.load_oracle()
    olib.orc_raw2ev.restype = C.POINTER(C.c_int)
    olib.orc_raw2ev.argtypes = [C.c_int]
    olib.orc_ev2raw.restype = C.POINTER(C.c_int)
    ev2raw = np.ctypeslib.as_array(olib.orc_ev2raw(), shape=(14 * EV,))
    top = np.ascontiguousarray(ev2raw[13 * EV:], dtype=np.uint16)

    def run(frames, black, white, coef, grid, segments=0, bad_xy=None, bad_vals=None):
        n = len(frames)
        h, w = frames[0].shape
        packed = np.stack([synth.pack_bits(f) for f in frames]).view(np.uint8)
        stride = packed.shape[1]
        assert stride % 16 == 0
        raw2ev = np.ascontiguousarray(np.ctypeslib.as_array(olib.orc_raw2ev(black), shape=(16384,)))
        out = np.full((n, h, w), 0xDEAD, np.uint16)
        assert out.ctypes.data % 16 == 0
        c8 = None if coef is None else np.ascontiguousarray(coef, np.int32)
        nb = 0 if bad_xy is None else len(bad_xy)
        xy = None if nb == 0 else np.ascontiguousarray(bad_xy, np.int32)
        bv = None if nb == 0 else np.ascontiguousarray(bad_vals, np.uint16)
        rc = lib.wide_emu_run(packed.ctypes.data, stride, out.ctypes.data, h * w, w, h, black, white, n, raw2ev.ctypes.data,
                              top.ctypes.data, None if c8 is None else c8.ctypes.data, grid, segments,
                              None if nb == 0 else xy.ctypes.data, nb, None if nb == 0 else bv.ctypes.data)
        return rc, out

    return run


def _want(oracle, frames, black, white, stripes):
    hh, ww = frames[0].shape
    want, state = oracle.single_iso_chain(frames, black, white, hh * ww * 14 // 8, chroma_smooth_method=3, fix_bad_pixels=0,
                                          fix_stripes=int(stripes))
    coef = None
    if stripes:
        needed, coef = state["stripes"]
        assert needed, "the synthetic column gains must need a correction"
    return want, coef


@pytest.mark.parametrize("w,h,n,stripes,grid,segments", [
    (640, 64, 3, True, 7, 0),        # 3 strips over 7 warps: strips own 3 / 2 / 2 warps, runs cross frame ends
    (640, 64, 3, True, 5, 3),        # the equal-segment split of the same frames
    (320, 38, 2, False, 9, 0),       # no stripe correction, 2 strips, runs of 4-5 rows
    (256, 24, 5, True, 64, 0),       # more warps than rows per strip allow: single-row pieces and idle warps
    (1920, 24, 1, True, 8, 0),       # the headline width: 8 strips, one warp each
])
def test_wide_kernel_source_on_host_matches_oracle(emu, oracle, w, h, n, stripes, grid, segments):
    black, white = 2048, 15000
    frames = [synth.make_frame(w, h, i, hot_cold=False, stripes=stripes) for i in range(n)]
    want, coef = _want(oracle, frames, black, white, stripes)
    rc, got = emu(frames, black, white, coef, grid, segments)
    assert rc == (2 if stripes else 0)
    for i in range(n):
        assert np.array_equal(got[i], want[i]), (i, int(np.count_nonzero(got[i] != want[i])))


def test_general_gains_dark_samples_and_other_levels(emu, oracle):
    """Eight general gains (kernel variant 1: gains 0 and 1 not 1.0), another black / white level, strong noise with
    samples at and below black (INT_MIN entries of raw2ev, wrap-around differences) and clipped highlights."""
    black, white, w, h = 1024, 12000, 320, 40
    frames = [synth.make_frame(w, h, i, black=black, white=white, stripes=True, noise_amp=700) for i in range(2)]
    for f in frames:
        f[10:20, 100:140] = np.minimum(f[10:20, 100:140].astype(np.int64) * 3 + 9000, 16383)      # clipped highlights
        f[24:30, 200:260:3] = black                                                               # samples exactly at black
    assert (frames[0] < black).any() and (frames[0] == black).any() and (frames[0] > white).any()
    coef = [65011, 66002, 66322, 64946, 66060, 64684, 66519, 65077]
    hh, ww = frames[0].shape
    want = []
    for f in frames:
        img = oracle.chroma_smooth(f, black, 3)
        want.append(oracle.stripes_apply(img, black, white, 1, coef))
    rc, got = emu(frames, black, white, coef, 4)
    assert rc == 1
    for i in range(2):
        assert np.array_equal(got[i], want[i]), (i, int(np.count_nonzero(got[i] != want[i])))


def test_ineligible_shapes_are_refused(emu):
    fr = [np.full((8, 96), 3000, np.uint16)]
    assert emu(fr, 2048, 15000, None, 4)[0] == -1                       # width not a multiple of 64
    fr = [np.full((8, 128), 3000, np.uint16)]
    assert emu(fr, 2048, 15000, [65536, 65536, 1 << 18, 65536, 65536, 65536, 65536, 65536], 4)[0] == -2


@pytest.mark.parametrize("w,h,n,grid,segments,density", [(640, 48, 3, 7, 0, 2e-3), (640, 48, 2, 5, 2, 2e-3), (256, 40, 2, 11, 0, 2e-2)])
def test_repaired_pixels_are_patched_into_the_staged_rows(emu, oracle, w, h, n, grid, segments, density):
    """--bad-pix: the repaired samples are written into the staged stream bytes before extraction, in both strips of a
    seam, at the right row of a run whatever piece of the column a warp owns (dense lists: many rows carry patches)."""
    black, white = 2048, 15000
    frames = [synth.make_frame(w, h, i, hot_cold=True, stripes=True, bad_density=density) for i in range(n)]
    plist = oracle.badpix_detect(frames[0], black, 0)
    assert len(plist) >= 8
    fixed = [oracle.badpix_apply(f, black, plist) for f in frames]
    vals = np.stack([f[plist[:, 1], plist[:, 0]] for f in fixed])
    assert any((fixed[i] != frames[i]).any() for i in range(n))
    want, state = oracle.single_iso_chain(frames, black, white, h * w * 14 // 8, chroma_smooth_method=3, fix_bad_pixels=1,
                                          fix_stripes=1)
    assert np.array_equal(state["badpix"], plist)
    needed, coef = state["stripes"]
    assert needed
    rc, got = emu(frames, black, white, coef, grid, segments, bad_xy=plist, bad_vals=vals)
    assert rc == 2
    for i in range(n):
        assert np.array_equal(got[i], want[i]), (i, int(np.count_nonzero(got[i] != want[i])))


def test_random_shapes_grids_and_splits(emu, oracle):
    """Seeded sweep over widths, heights, frame counts, grid sizes (down to one warp per strip) and both work splits,
    with repaired pixels: every output pixel is written exactly once and equals the oracle's."""
    rng = np.random.default_rng(20261017)
    black, white = 2048, 15000
    for case in range(6):
        w = int(rng.choice([128, 192, 256, 320, 512, 704]))
        h = 2 * int(rng.integers(8, 26))
        n = int(rng.integers(1, 5))
        nstrips = -(-w // 240)
        grid = int(rng.integers(nstrips, 4 * nstrips + 6))
        segments = int(rng.integers(1, 4)) if case % 3 == 2 else 0
        frames = [synth.make_frame(w, h, 10 * case + i, hot_cold=True, stripes=True, bad_density=3e-3) for i in range(n)]
        want, state = oracle.single_iso_chain(frames, black, white, h * w * 14 // 8, chroma_smooth_method=3, fix_bad_pixels=1,
                                              fix_stripes=1)
        plist = state["badpix"]
        needed, coef = state["stripes"]
        vals = np.stack([oracle.badpix_apply(f, black, plist)[plist[:, 1], plist[:, 0]] for f in frames]) if len(plist) else None
        rc, got = emu(frames, black, white, coef if needed else None, grid, segments,
                      bad_xy=plist if len(plist) else None, bad_vals=vals)
        assert rc >= 0, (case, w, h, n, grid, segments)
        if not needed:                                     # coefficients inside 0.998 .. 1.002: stripes.c leaves the frame alone
            want = [oracle.chroma_smooth(oracle.badpix_apply(f, black, plist), black, 3) for f in frames]
        for i in range(n):
            assert np.array_equal(got[i], want[i]), (case, w, h, n, grid, segments, i, int(np.count_nonzero(got[i] != want[i])))
