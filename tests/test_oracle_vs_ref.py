"""Pins the CPU restatement (oracle/*.c) against the UNMODIFIED reference compiled into oracle/_ref.

The reference ships no tests or golden vectors for this path, so this comparison --
plus the fixtures in tests/golden that were generated from the same build -- is what anchors parity.
Skipped where oracle/_ref was never built (it needs /root/reference at build time).
"""
import ctypes as C
import os

import numpy as np
import pytest

from mlvfs_b200 import mlvformat as F, synth


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_luts_match_reference(oracle, ref):
    orc = oracle.load_oracle()
    for black in (0, 2048, 1024, 16384):
        n = 16384 + black
        r = np.ctypeslib.as_array(ref.get_raw2ev(black), (n,))
        o = np.ctypeslib.as_array(orc.orc_raw2ev(black), (n,))
        assert np.array_equal(r, o)
        rf = np.ctypeslib.as_array(ref.get_raw2evf(black), (n,))
        of = np.ctypeslib.as_array(orc.orc_raw2evf(black), (n,))
        assert np.array_equal(rf[black + 1:], of[black + 1:])
    lo = 10 * 32768
    r = np.ctypeslib.as_array(C.cast(C.addressof(ref.get_ev2raw().contents) - 4 * lo, C.POINTER(C.c_int)), (24 * 32768,))
    o = np.ctypeslib.as_array(C.cast(C.addressof(orc.orc_ev2raw().contents) - 4 * lo, C.POINTER(C.c_int)), (24 * 32768,))
    assert np.array_equal(r, o)
    assert orc.orc_raw2ev(16385) is None or not orc.orc_raw2ev(16385)


def test_rand_matches_libc(oracle):
    orc = oracle.load_oracle()
    st = oracle.RandState()
    orc.orc_rand_seed(C.byref(st), 1)
    libc = C.CDLL(None)
    libc.srand(1)
    assert all(libc.rand() == orc.orc_rand_next(C.byref(st)) for _ in range(200000))
    orc.orc_rand_seed(C.byref(st), 777)
    libc.srand(777)
    assert all(libc.rand() == orc.orc_rand_next(C.byref(st)) for _ in range(1000))


@pytest.mark.parametrize("bpp", [8, 10, 12, 14])
def test_unpack_matches_reference(oracle, ref, bpp):
    w, h = 328, 66
    hdr = F.make_frame_headers(w, h, bpp=bpp)
    rng = np.random.default_rng(bpp)
    img = rng.integers(0, 1 << bpp, size=(h, w), dtype=np.uint16)
    words = np.concatenate([synth.pack_bits(img, bpp), np.zeros(2, np.uint16)])
    out = np.zeros(w * h, np.uint16)
    ref.dng_get_image_data(C.byref(hdr), _p(words), _p(out), 0, out.nbytes)
    assert np.array_equal(out, img.ravel())
    assert np.array_equal(oracle.unpack(words, w * h, bpp), out)
    for offset, size in [(2, 1000), (4096, 8192), (w * h * 2 - 512, 512)]:
        first_word = (offset // 2) * bpp // 16
        sub = np.zeros(size // 2, np.uint16)
        ref.dng_get_image_data(C.byref(hdr), _p(words[first_word:]), _p(sub), offset, size)
        assert np.array_equal(oracle.unpack(words, w * h, bpp, offset=offset, nbytes=size), sub)


@pytest.mark.parametrize("method", [2, 3, 5])
def test_chroma_smooth_matches_reference(oracle, ref, method):
    w, h = 320, 180
    hdr = F.make_frame_headers(w, h)
    rng = np.random.default_rng(3)
    for k, img in enumerate([synth.make_frame(w, h, 1), rng.integers(2040, 2070, (h, w)).astype(np.uint16)]):
        if k == 1:
            img[::3, ::5] = 2048          # INT_MIN LUT entries -> wrap-around arithmetic
            img[40:90, 60:200] += 2500
        want = img.copy()
        ref.chroma_smooth(C.byref(hdr), _p(want), method)
        assert np.array_equal(oracle.chroma_smooth(img, 2048, method), want)


@pytest.mark.parametrize("aggressive", [0, 1])
def test_bad_pixels_match_reference(oracle, ref, aggressive):
    w, h = 480, 270
    img = synth.make_frame(w, h, 0, hot_cold=True, bad_density=3e-4)
    rng = np.random.default_rng(5)
    for _ in range(40):                      # clusters -> order-dependent repairs
        x0, y0 = int(rng.integers(8, w - 24)), int(rng.integers(8, h - 24))
        for k in range(int(rng.integers(2, 6))):
            img[y0 + (2 * k if k % 2 else 0), x0 + (0 if k % 2 else 2 * k)] = 16000
    for dual in (0, 1):
        hdr = F.make_frame_headers(w, h, file_guid=0x1000 + aggressive * 2 + dual)
        want = img.copy()
        with oracle.quiet_stdout():
            ref.fix_bad_pixels(C.byref(hdr), _p(want), aggressive, dual)
        lst = oracle.badpix_detect(img, 2048, aggressive)
        assert len(lst) > 50
        assert np.array_equal(oracle.badpix_apply(img, 2048, lst, dual_iso=dual), want)


def test_focus_pixels_match_reference(oracle, ref, tmp_path):
    """fix_focus_pixels with a synthetic .fpm map in CWD, including border entries (cs.c:479-500)."""
    w, h = 400, 200
    cam = 0x80000326
    hdr = F.make_frame_headers(w, h, camera_model=cam, raw_width=1808, raw_height=727, pan_x=16, pan_y=10)
    crop = ((16 + 7) & ~7, 10 & ~1)
    rng = np.random.default_rng(9)
    pts = [(int(x) + crop[0], int(y) + crop[1]) for x, y in zip(rng.integers(-2, w + 2, 600), rng.integers(-2, h + 2, 600))]
    pts += [(x + crop[0], 50 + crop[1]) for x in range(100, 130, 2)]           # a dependent run
    pts += [(0 + crop[0], 5 + crop[1]), (w - 1 + crop[0], 7 + crop[1]), (3 + crop[0], 1 + crop[1]), (200 + crop[0], h - 1 + crop[1])]
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with open("%x_%dx%d.fpm" % (cam, 1808, 727), "w") as f:
            for x, y in pts:
                f.write(f"{x} \t {y}\n")
        img = synth.make_frame(w, h, 4)
        for dual in (0, 1):
            want = img.copy()
            with oracle.quiet_stdout():
                ref.fix_focus_pixels(C.byref(hdr), _p(want), dual)
            got = oracle.focuspix_apply(img, 2048, np.array(pts, np.int32), crop=crop, dual_iso=dual)
            assert np.array_equal(got, want)
            assert (want != img).sum() > 300
    finally:
        os.chdir(cwd)


def test_stripes_match_reference(oracle, ref):
    w, h = 1920, 1080
    hdr = F.make_frame_headers(w, h)
    ri = hdr.rawi_hdr.raw_info
    img = synth.make_frame(w, h, 0, stripes=True)
    C.CDLL(None).srand(1)
    ref.stripes_new_correction.argtypes = [C.c_char_p]
    corr = ref.stripes_new_correction(b"/tmp/oracle_stripes_test.MLV")

    class Corr(C.Structure):
        _fields_ = [("next", C.c_void_p), ("name", C.c_char_p), ("needed", C.c_int), ("coef", C.c_int * 8)]

    ref.stripes_compute_correction.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_size_t]
    ref.stripes_apply_correction.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_size_t]
    ref.stripes_compute_correction(C.byref(hdr), corr, _p(img), 0, img.size)
    c = Corr.from_address(corr)
    needed, coef = oracle.stripes_compute(img, ri.black_level, ri.white_level, ri.frame_size)
    assert needed == c.needed == 1
    assert list(coef) == list(c.coef)
    want = img.copy()
    ref.stripes_apply_correction(C.byref(hdr), corr, _p(want), 0, want.size)
    assert np.array_equal(oracle.stripes_apply(img, ri.black_level, ri.white_level, needed, coef), want)


def test_process_frame_chain_matches_reference(oracle, ref, tmp_path):
    """Whole C2 chain through the reference's own process_frame (main.c:908) on a synthetic MLV."""
    w, h = 1920, 1080
    clip = str(tmp_path / "C2.MLV")
    hdr, frames = synth.make_clip(clip, w, h, 3, variant=dict(hot_cold=True, stripes=True))
    ri = hdr.rawi_hdr.raw_info
    ref.ref_set_mlv_dir(str(tmp_path).encode())
    ref.ref_set_options(3, 1, 1, 0, 0, 0, 0, 0, 0)
    C.CDLL(None).srand(1)
    want = []
    with oracle.quiet_stdout():
        for i in range(3):
            out = np.zeros((h, w), np.uint16)
            assert ref.ref_process_frame(b"/C2.MLV/C2_%06d.dng" % i, _p(out), out.nbytes, None) == out.nbytes
            want.append(out)
    got, state = oracle.single_iso_chain(frames, ri.black_level, ri.white_level, ri.frame_size,
                                         chroma_smooth_method=3, fix_bad_pixels=1, fix_stripes=1)
    assert len(state["badpix"]) > 10 and state["stripes"][0] == 1
    for i in range(3):
        assert np.array_equal(got[i], want[i]), i


def test_pattern_noise_matches_reference(oracle, ref):
    for (w, h) in [(320, 180), (130, 256)]:
        img = synth.make_frame(w, h, 0)
        rng = np.random.default_rng(1)
        img = (img.astype(np.int32) + rng.integers(-12, 13, size=(1, w)) + rng.integers(-9, 10, size=(h, 1)))
        img = img.clip(0, 16383).astype(np.uint16)
        img[20:40, 50:90] = 15200
        want = img.copy()
        with oracle.quiet_stdout():
            ref.fix_pattern_noise(_p(want), w, h, 15000, 0)
        assert np.array_equal(oracle.fix_pattern_noise(img, 15000), want)


def test_lj92_codec_matches_reference(oracle, ref):
    """Our decoder reads the reference encoder's stream; the reference decoder reads our encoder's stream."""
    w, h = 352, 198
    img = synth.make_frame(w, h, 2, hot_cold=True, bad_density=1e-3)
    tiled = np.ascontiguousarray(synth.quadrant_interleave(img))
    # reference encode -> oracle decode
    enc, n = C.POINTER(C.c_uint8)(), C.c_int()
    ref.lj92_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p]
    assert ref.lj92_encode(_p(tiled), w, h, 14, w * h, 0, None, 0, C.byref(enc), C.byref(n)) == 0
    stream = np.ctypeslib.as_array(enc, (n.value,)).copy()
    payload = np.concatenate([np.array([w * h * 2], dtype="<u4").view(np.uint8), stream])
    assert np.array_equal(oracle.lj92_decode_payload(payload, w, h), img)
    # oracle encode -> reference decode
    mine = oracle.lj92_encode(tiled)
    hd, W, H, B = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
    assert ref.lj92_open(C.byref(hd), _p(mine), int(mine.size), C.byref(W), C.byref(H), C.byref(B)) == 0
    assert (W.value, H.value, B.value) == (w, h, 14)
    out = np.zeros(w * h, np.uint16)
    ref.lj92_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    assert ref.lj92_decode(hd, _p(out), w * h, 0, None, 0) == 0
    assert np.array_equal(out.reshape(h, w), tiled)


def _ref_cr2hdr20(ref, oracle, hdr, img, interp, fullres, alias, cs, badpix):
    ref.cr2hdr20_convert_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    out = img.copy()
    with oracle.quiet_stdout():
        r = ref.cr2hdr20_convert_data(C.byref(hdr), _p(out), interp, fullres, alias, cs, badpix)
    return r, out


@pytest.mark.parametrize("w,h,cs,alias,badpix,fullres", [
    (640, 360, 0, 0, 0, 1), (640, 360, 5, 1, 0, 1), (640, 362, 3, 1, 2, 1), (320, 180, 2, 0, 1, 1), (384, 216, 3, 1, 0, 0)])
def test_dual_iso_mean23_matches_reference(oracle, ref, w, h, cs, alias, badpix, fullres):
    """cr2hdr20_convert_data with --mean23: statistics, exposure match, interpolation, alias map, blend."""
    img = synth.make_frame(w, h, 0, dual_iso=True, hot_cold=True, bad_density=1e-4)
    hdr = F.make_frame_headers(w, h, file_guid=0x9100 + cs * 64 + alias * 16 + badpix * 4 + fullres)
    r, want = _ref_cr2hdr20(ref, oracle, hdr, img, 1, fullres, alias, cs, badpix)
    rc, got, info = oracle.cr2hdr20(img, 2048, 15000, interp_method=1, fullres=fullres, use_alias_map=alias,
                                    chroma_smooth_method=cs, fix_bad_pixels_mode=badpix)
    assert r == rc == 1
    assert hdr.rawi_hdr.raw_info.black_level == 2048 * 4 and hdr.rawi_hdr.raw_info.white_level == 15000 * 4
    assert list(info.is_bright) == [0, 0, 1, 1] and 2.9 < info.corr_ev < 3.1
    assert np.array_equal(got, want)


def test_dual_iso_other_row_phase_and_gbrg(oracle, ref):
    """Bright rows at y%4 in {0,1} (pattern BBdd) and a GBRG frame (one row cropped, hdr.c:1784-1791)."""
    w, h = 480, 272
    base = synth.make_frame(w, h + 3, 3, dual_iso=True)
    for k, img in enumerate([base[2:h + 2], base[1:h + 1], base[3:h + 3]]):
        img = np.ascontiguousarray(img)
        hdr = F.make_frame_headers(w, h, file_guid=0x9200 + k)
        r, want = _ref_cr2hdr20(ref, oracle, hdr, img, 1, 1, 1, 3, 0)
        rc, got, info = oracle.cr2hdr20(img, 2048, 15000, interp_method=1, chroma_smooth_method=3)
        assert r == rc, k
        assert np.array_equal(got, want), k


def test_dual_iso_rejects_plain_footage(oracle, ref):
    w, h = 320, 180
    img = synth.make_frame(w, h, 1, hot_cold=True, bad_density=1e-4)
    hdr = F.make_frame_headers(w, h, file_guid=0x9300)
    r, want = _ref_cr2hdr20(ref, oracle, hdr, img, 1, 1, 1, 0, 1)
    rc, got, _ = oracle.cr2hdr20(img, 2048, 15000, interp_method=1, fix_bad_pixels_mode=1)
    assert r == rc == 0
    assert hdr.rawi_hdr.raw_info.black_level == 2048
    assert np.array_equal(got, want)          # still carries the horizontal bad-pixel repairs (hdr.c:1944-1948)


@pytest.mark.parametrize("w,h,shift,white", [(640, 360, 0, 15000), (642, 362, 1, 15000), (640, 360, 2, 12000), (640, 360, 3, 15000)])
def test_dual_iso_preview_matches_reference(oracle, ref, w, h, shift, white):
    base = synth.make_frame(w, h + 4, 0, dual_iso=True, white=white)
    img = np.ascontiguousarray(base[shift:shift + h])
    if white == 12000:
        img[100:140, 200:300] = 0                 # deep shadows on dark rows: patched from the bright rows
    hdr = F.make_frame_headers(w, h, white=white)
    want = img.copy()
    ref.hdr_convert_data.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_size_t]
    assert ref.hdr_convert_data(C.byref(hdr), _p(want), 0, want.nbytes) == 1
    rc, got = oracle.hdr_preview(img, 2048, white)
    assert rc == 1 and np.array_equal(got, want)
    assert hdr.rawi_hdr.raw_info.black_level == 8192
    plain = synth.make_frame(w, h, 1)
    assert oracle.hdr_preview(plain, 2048, white)[0] == 0


def test_deflicker_matches_reference(oracle, ref):
    w, h = 640, 360
    for seed, target in [(0, 3000), (3, 5000), (5, 2100)]:
        img = synth.make_frame(w, h, seed)
        hdr = F.make_frame_headers(w, h)
        ref.ref_deflicker.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        ref.ref_deflicker(C.byref(hdr), target, _p(img), img.nbytes)
        want = (hdr.rawi_hdr.raw_info.exposure_bias[0], hdr.rawi_hdr.raw_info.exposure_bias[1])
        assert oracle.deflicker(img, 14, 2048, target) == want
    flat = np.full((h, w), 3000, np.uint16)        # > 65535 samples in one bin: the uint16 counter wraps
    hdr = F.make_frame_headers(w, h)
    ref.ref_deflicker(C.byref(hdr), 4000, _p(flat), flat.nbytes)
    assert oracle.deflicker(flat, 14, 2048, 4000) == (hdr.rawi_hdr.raw_info.exposure_bias[0], 10000)


# ---- AMaZE + edge-directed dual ISO (amaze_demosaic_RT.c, hdr.c:917-1229) -------------------------------

@pytest.mark.parametrize("w,h", [(160, 160), (288, 200), (256, 344), (384, 270), (128, 96)])
def test_amaze_planes_match_reference(oracle, ref, w, h):
    """The SSE2 build of amaze_demosaic_RT vs the lane-by-lane restatement: float planes bit for bit,
    including partial right/bottom tiles and the reference's border-fill overrun (h - top in (144, 160))."""
    raw = synth.amaze_test_mosaic(w, h, w * 7 + h)
    want = oracle.ref_amaze_demosaic(raw)
    got = oracle.amaze_demosaic(raw)
    for name, a, b in zip("RGB", got, want):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), name


@pytest.mark.parametrize("w,h", [(256, 216), (384, 328), (512, 136)])
def test_amaze_tiles_are_independent(oracle, w, h):
    """What lets the CUDA path run one block per tile: with width % 128 == 0 (all BASELINE configs) the
    reference's tile-to-tile persistent work planes never influence the output.  `fresh_tiles=1` clears the
    block per tile, `-1` poisons every plane except pmwt (zeroed) with NaN."""
    raw = synth.amaze_test_mosaic(w, h, 3)
    base = oracle.amaze_demosaic(raw, fresh_tiles=0)
    for mode in (1, -1):
        other = oracle.amaze_demosaic(raw, fresh_tiles=mode)
        for a, b in zip(base, other):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), mode


@pytest.mark.parametrize("w,h,cs,alias,badpix,fullres", [
    (640, 360, 0, 0, 0, 1), (640, 362, 3, 1, 2, 1), (384, 216, 5, 1, 0, 0), (300, 200, 0, 1, 1, 1)])
def test_dual_iso_amaze_matches_reference(oracle, ref, w, h, cs, alias, badpix, fullres):
    """cr2hdr20_convert_data with --amaze-edge (interp_method 0): squeeze, AMaZE, gray, edge-direction search,
    edge-directed interpolation, then the shared mix/alias/blend stages."""
    img = synth.make_frame(w, h, 0, dual_iso=True, hot_cold=True, bad_density=1e-4)
    hdr = F.make_frame_headers(w, h, file_guid=0xB100 + cs * 64 + alias * 16 + badpix * 4 + fullres + w)
    r, want = _ref_cr2hdr20(ref, oracle, hdr, img, 0, fullres, alias, cs, badpix)
    rc, got, info = oracle.cr2hdr20(img, 2048, 15000, interp_method=0, fullres=fullres, use_alias_map=alias,
                                    chroma_smooth_method=cs, fix_bad_pixels_mode=badpix)
    assert r == rc == 1
    assert np.array_equal(got, want)
