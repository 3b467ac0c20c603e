"""GPU parity: single-ISO chain (unpack, chroma smoothing, bad pixels, stripes) vs the CPU oracle.

Bit-exact for every stage (integer / LUT arithmetic).  Calls go through the C ABI
(libmlvfs_b200.so) exactly as the host frame builder makes them.
"""
import ctypes as C

import numpy as np
import pytest

import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F, synth

pytestmark = pytest.mark.gpu

SIZES = [(1920, 1080), (256, 130), (72, 40)]


def _hdr(w, h, **kw):
    return F.make_frame_headers(w, h, **kw)


@pytest.mark.parametrize("w,h", SIZES)
def test_unpack14_matches_oracle(fresh_ctx, oracle, w, h):
    hdr = _hdr(w, h)
    img = synth.make_frame(w, h, 3)
    words = synth.pack_bits(img)
    want = oracle.unpack(words, w * h)
    assert (want == img.ravel()).all()
    out, res = fresh_ctx.process_frame(hdr, words, M.Options(), "unpack.MLV")
    assert res.status == 0 and not res.is_dual_iso
    assert np.array_equal(out.ravel(), want)
    # drop-in entry point (dng.h:31)
    n, raw = M.dng_get_image_data(hdr, words)
    assert n == w * h * 2
    assert np.array_equal(raw.view(np.uint16), want)


@pytest.mark.parametrize("bpp", [8, 10, 12, 14])
def test_unpack_other_depths(fresh_ctx, oracle, bpp):
    w, h = 328, 66      # npix not a multiple of 1024 groups; exercises the tail paths
    hdr = _hdr(w, h, bpp=bpp, black=2048 >> (14 - bpp), white=15000 >> (14 - bpp))
    rng = np.random.default_rng(bpp)
    img = rng.integers(0, 1 << bpp, size=(h, w), dtype=np.uint16)
    words = synth.pack_bits(img, bpp)
    want = oracle.unpack(words, w * h, bpp)
    assert (want == img.ravel()).all()
    n, raw = M.dng_get_image_data(hdr, words)
    assert np.array_equal(raw.view(np.uint16), want)


def test_unpack_subrange_offsets(fresh_ctx, oracle):
    """Pismo-style ranged reads (win/mlvfs-pfm.cpp:1221-1237): even byte offsets into the frame."""
    w, h = 640, 48
    hdr = _hdr(w, h)
    img = synth.make_frame(w, h, 1)
    words = np.concatenate([synth.pack_bits(img), np.zeros(2, np.uint16)])
    for offset, size in [(0, 4096), (2, 1000), (4096, 8192), (w * h * 2 - 512, 512), (32768, 16384)]:
        first_word = (offset // 2) * 14 // 16
        want = oracle.unpack(words, w * h, 14, offset=offset, nbytes=size)
        n, raw = M.dng_get_image_data(hdr, words[first_word:], offset=offset, max_size=size)
        assert n == size
        assert np.array_equal(raw.view(np.uint16), want), (offset, size)


@pytest.mark.parametrize("method", [2, 3, 5])
@pytest.mark.parametrize("w,h", [(1920, 1080), (256, 130), (70, 37)])
def test_chroma_smooth_matches_oracle(fresh_ctx, oracle, method, w, h):
    hdr = _hdr(w, h)
    img = synth.make_frame(w, h, 5)
    want = oracle.chroma_smooth(img, 2048, method)
    got = M.chroma_smooth(hdr, img.copy(), method)
    assert np.array_equal(got, want)
    assert (want != img).any() or min(w, h) < 12


def test_chroma_smooth_black_pixels_wrap(fresh_ctx, oracle):
    """Pixels exactly at black hit the INT_MIN LUT entry: wrap-around arithmetic must match (A.2)."""
    w, h = 128, 64
    hdr = _hdr(w, h)
    rng = np.random.default_rng(7)
    img = rng.integers(2040, 2060, size=(h, w), dtype=np.uint16)
    img[::3, ::5] = 2048
    img[10:30, 20:60] += 3000
    for method in (2, 3, 5):
        assert np.array_equal(M.chroma_smooth(hdr, img.copy(), method), oracle.chroma_smooth(img, 2048, method))


@pytest.mark.parametrize("aggressive", [0, 1])
def test_bad_pixels_detect_and_fix(fresh_ctx, oracle, aggressive):
    w, h = 1920, 1080
    hdr = _hdr(w, h)
    img = synth.make_frame(w, h, 0, hot_cold=True, bad_density=2e-5)
    want_list = oracle.badpix_detect(img, 2048, aggressive)
    want = oracle.badpix_apply(img, 2048, want_list)
    got = M.fix_bad_pixels(hdr, img.copy(), aggressive, 0)
    got_list = M.Context.default().get_bad_pixels(hdr.file_hdr.fileGuid, aggressive)
    assert len(want_list) > 20
    assert np.array_equal(got_list, want_list)          # same entries, same (raster) order
    assert np.array_equal(got, want)
    # second frame of the same clip reuses the map (no re-detection), dual-ISO flag picks the 1-D interpolator
    img2 = synth.make_frame(w, h, 1, hot_cold=True, bad_density=2e-5)
    assert np.array_equal(M.fix_bad_pixels(hdr, img2.copy(), aggressive, 1),
                          oracle.badpix_apply(img2, 2048, want_list, dual_iso=1))


def test_bad_pixel_clusters_are_order_exact(fresh_ctx, oracle):
    """Adjacent defects read each other's repaired values (A.4): exercise deep dependency levels."""
    w, h = 256, 128
    hdr = _hdr(w, h, file_guid=0xABCDEF01)
    img = synth.make_frame(w, h, 2)
    rng = np.random.default_rng(11)
    # clusters: runs along rows/columns at distance 1, 2, 3 so that chains form
    for _ in range(60):
        x0, y0 = int(rng.integers(8, w - 24)), int(rng.integers(8, h - 24))
        for k in range(int(rng.integers(2, 7))):
            if rng.integers(0, 2):
                img[y0, x0 + 2 * k] = 16000
            else:
                img[y0 + 2 * k, x0] = 16000
    want_list = oracle.badpix_detect(img, 2048, 1)
    assert len(want_list) > 100
    want = oracle.badpix_apply(img, 2048, want_list)
    got = M.fix_bad_pixels(hdr, img.copy(), 1, 0)
    assert np.array_equal(got, want)


def test_stripes_compute_and_apply(fresh_ctx, oracle):
    w, h = 1920, 1080
    hdr = _hdr(w, h)
    ri = hdr.rawi_hdr.raw_info
    img = synth.make_frame(w, h, 0, stripes=True)
    needed, coef = oracle.stripes_compute(img, ri.black_level, ri.white_level, ri.frame_size)
    assert needed == 1
    corr = M.stripes_compute_correction(hdr, img, "stripes_clip.MLV")
    got_coef = np.array(list(corr.contents.coeffficients))
    # tolerance stage (A.5): same dither stream -> normally identical; allow one histogram bin (2^-15 EV)
    assert corr.contents.correction_needed == 1
    assert np.max(np.abs(got_coef - coef)) <= 2, (got_coef, coef)
    want = oracle.stripes_apply(img, ri.black_level, ri.white_level, needed, got_coef)
    got = M.stripes_apply_correction(hdr, corr, img.copy())
    assert np.array_equal(got, want)


@pytest.mark.parametrize("opts", [
    dict(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=1),      # BASELINE config 2
    dict(chroma_smooth=2, fix_bad_pixels=2, fix_stripes=0),
    dict(chroma_smooth=5, fix_bad_pixels=0, fix_stripes=1),
    dict(chroma_smooth=0, fix_bad_pixels=1, fix_stripes=1),
    dict(),                                                      # BASELINE config 1
])
def test_full_single_iso_chain(fresh_ctx, oracle, opts):
    """process_frame order (main.c:966-997) over several frames; per-clip state comes from frame 0."""
    w, h = 1920, 1080
    hdr = _hdr(w, h)
    ri = hdr.rawi_hdr.raw_info
    frames = [synth.make_frame(w, h, i, hot_cold=True, stripes=True) for i in range(3)]
    want, state = oracle.single_iso_chain(frames, ri.black_level, ri.white_level, ri.frame_size,
                                          chroma_smooth_method=opts.get("chroma_smooth", 0),
                                          fix_bad_pixels=opts.get("fix_bad_pixels", 0),
                                          fix_stripes=opts.get("fix_stripes", 0))
    o = M.Options(**opts)
    for i, fr in enumerate(frames):
        out, res = fresh_ctx.process_frame(hdr, synth.pack_bits(fr), o, "chain.MLV")
        assert res.status == 0
        if opts.get("fix_stripes"):
            needed, coef = fresh_ctx.get_stripes("chain.MLV")
            assert needed == state["stripes"][0]
            assert np.array_equal(coef, state["stripes"][1]), "stripe coefficients differ from the oracle"
        d = out.astype(np.int32) - want[i].astype(np.int32)
        assert np.abs(d).max() == 0, f"frame {i}: {np.count_nonzero(d)} px differ, max {np.abs(d).max()}"


def test_batch_device_matches_per_frame(fresh_ctx, oracle):
    """The device-resident batch entry (what bench.py times) gives the same frames as the per-frame call."""
    torch = pytest.importorskip("torch")
    w, h, n = 1920, 1080, 6
    hdr = _hdr(w, h)
    ri = hdr.rawi_hdr.raw_info
    frames = [synth.make_frame(w, h, i, hot_cold=True, stripes=True) for i in range(n)]
    want, _ = oracle.single_iso_chain(frames, ri.black_level, ri.white_level, ri.frame_size,
                                      chroma_smooth_method=3, fix_bad_pixels=1, fix_stripes=1)
    packed = np.stack([synth.pack_bits(f) for f in frames])
    stride = packed.shape[1] * 2
    assert stride % 16 == 0
    d_in = torch.from_numpy(packed.view(np.int16)).cuda()
    d_out = torch.empty((n, h * w), dtype=torch.int16, device="cuda")
    o = M.Options(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=1)
    for rep in range(2):        # second pass runs with cached per-clip state (fused stripes epilogue)
        fresh_ctx.process_batch_device(hdr, o, "batch.MLV", d_in.data_ptr(), stride, stride, d_out.data_ptr(), h * w, n,
                                       torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(np.uint16).reshape(n, h, w)
        for i in range(n):
            assert np.array_equal(got[i], want[i]), (rep, i)


@pytest.mark.parametrize("w,h,stripes", [(1920, 1080, 1), (656, 362, 1), (1040, 94, 0)])
def test_fused_kernel_odd_geometries(fresh_ctx, oracle, w, h, stripes):
    """The steady-state C2 chain runs in one fused kernel (fused.cu); it must give the oracle's frames also at
    segment / strip edges of odd geometries and with a dense bad-pixel list."""
    hdr = _hdr(w, h)
    ri = hdr.rawi_hdr.raw_info
    frames = [synth.make_frame(w, h, i, hot_cold=True, stripes=bool(stripes), bad_density=3e-4) for i in range(4)]
    want, _ = oracle.single_iso_chain(frames, ri.black_level, ri.white_level, ri.frame_size,
                                      chroma_smooth_method=3, fix_bad_pixels=1, fix_stripes=stripes)
    o = M.Options(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=stripes)
    name = f"fusedgeo_{w}x{h}.MLV"
    for i, fr in enumerate(frames):              # frames 0, 1 build the per-clip state / take the general path
        out, res = fresh_ctx.process_frame(hdr, synth.pack_bits(fr), o, name)
        assert res.status == 0
        assert np.array_equal(out, want[i]), (i, int(np.count_nonzero(out != want[i])))


@pytest.mark.parametrize("w,h,n,stripes,density", [(1920, 1080, 8, 1, 3e-5), (640, 362, 5, 1, 3e-4), (1088, 94, 3, 0, 3e-4),
                                                   (3840, 160, 3, 1, 1e-4)])
def test_wide_fused_kernel_batches(fresh_ctx, oracle, monkeypatch, w, h, n, stripes, density):
    """Large device batches of the C2 chain take the persistent wide kernel (fused_wide.cuh: shared-memory EV
    tables, 8 quad columns per lane, cp.async row staging, patches written into the staged bytes).  Same frames
    as the oracle, bit for bit, including strip / segment seams, image borders and a dense bad-pixel list."""
    torch = pytest.importorskip("torch")
    monkeypatch.setenv("MLVB_WIDE_MIN_ROWS", "1")
    hdr = _hdr(w, h)
    ri = hdr.rawi_hdr.raw_info
    frames = [synth.make_frame(w, h, i, hot_cold=True, stripes=bool(stripes), bad_density=density) for i in range(n)]
    want, _ = oracle.single_iso_chain(frames, ri.black_level, ri.white_level, ri.frame_size,
                                      chroma_smooth_method=3, fix_bad_pixels=1, fix_stripes=stripes)
    packed = np.stack([synth.pack_bits(f) for f in frames])
    stride = packed.shape[1] * 2
    assert stride % 16 == 0
    d_in = torch.from_numpy(packed.view(np.int16)).cuda()
    d_out = torch.empty((n, h * w), dtype=torch.int16, device="cuda")
    o = M.Options(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=stripes)
    name = f"wide_{w}x{h}.MLV"
    for rep in range(3):        # pass 0 builds the per-clip state on the general path
        d_out.zero_()
        fresh_ctx.process_batch_device(hdr, o, name, d_in.data_ptr(), stride, stride, d_out.data_ptr(), h * w, n,
                                       torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(np.uint16).reshape(n, h, w)
        for i in range(n):
            assert np.array_equal(got[i], want[i]), (rep, i, int(np.count_nonzero(got[i] != want[i])),
                                                     np.argwhere(got[i] != want[i])[:8].tolist())
    assert fresh_ctx.path_count(1) >= 2, "the wide kernel did not run"


@pytest.mark.parametrize("black,white,bits_noise", [(1024, 16000, 16), (4000, 14000, 64), (2048, 15000, 600)])
def test_wide_fused_kernel_levels_and_dark_frames(fresh_ctx, oracle, monkeypatch, black, white, bits_noise):
    """Other black / white levels (the shared-memory raw2ev copy is cut for the clip's black level) and frames with
    many samples at or below black: raw2ev[black] is INT_MIN and the EV arithmetic wraps exactly like the reference's
    (chroma_smooth.c:32); samples above white are clamped by the stripe stage."""
    torch = pytest.importorskip("torch")
    monkeypatch.setenv("MLVB_WIDE_MIN_ROWS", "1")
    w, h, n = 704, 258, 4
    hdr = _hdr(w, h, black=black, white=white, file_guid=0xC0DE0000 + black)
    frames = []
    for i in range(n):
        f = synth.make_frame(w, h, i, black=black, white=white, hot_cold=True, stripes=True, noise_amp=bits_noise, bad_density=1e-4)
        rng = np.random.default_rng(100 + i)
        dark = rng.random((h, w)) < 0.05
        f[dark] = np.clip(black + rng.integers(-40, 3, size=int(dark.sum())), 0, 16383)     # around and exactly at black
        frames.append(f)
    want, _ = oracle.single_iso_chain(frames, black, white, hdr.rawi_hdr.raw_info.frame_size,
                                      chroma_smooth_method=3, fix_bad_pixels=1, fix_stripes=1)
    packed = np.stack([synth.pack_bits(f) for f in frames])
    stride = packed.shape[1] * 2
    d_in = torch.from_numpy(packed.view(np.int16)).cuda()
    d_out = torch.empty((n, h * w), dtype=torch.int16, device="cuda")
    o = M.Options(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=1)
    name = f"widelvl_{black}.MLV"
    for rep in range(3):
        d_out.zero_()
        fresh_ctx.process_batch_device(hdr, o, name, d_in.data_ptr(), stride, stride, d_out.data_ptr(), h * w, n,
                                       torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(np.uint16).reshape(n, h, w)
        for i in range(n):
            assert np.array_equal(got[i], want[i]), (rep, i, int(np.count_nonzero(got[i] != want[i])),
                                                     np.argwhere(got[i] != want[i])[:8].tolist())
    if bits_noise < 100:        # the very noisy clip has chained bad pixels: level-scheduled general path by design
        assert fresh_ctx.path_count(1) >= 2, "the wide kernel did not run"
