"""Oracle restatement vs the committed golden fixtures (tests/golden/single_iso.npz, produced by
tests/golden/make_golden.py from the unmodified reference).  Runs anywhere, no /root/reference needed."""
import ctypes as C
import os

import numpy as np
import pytest

from mlvfs_b200 import mlvformat as F, synth

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "single_iso.npz"))
W, H = 192, 96


def test_unpack_golden(oracle):
    for bpp in (8, 10, 12, 14):
        assert np.array_equal(oracle.unpack(G[f"unpack{bpp}_words"], W * H, bpp), G[f"unpack{bpp}_out"])


def test_chroma_smooth_golden(oracle):
    img = synth.make_frame(W, H, 5)
    for m in (2, 3, 5):
        assert np.array_equal(oracle.chroma_smooth(img, 2048, m), G[f"cs{m}_out"])


def test_bad_pixels_golden(oracle):
    bp = G["badpix_in"]
    for aggr in (0, 1):
        lst = oracle.badpix_detect(bp, 2048, aggr)
        assert np.array_equal(oracle.badpix_apply(bp, 2048, lst), G[f"badpix{aggr}_out"])


def test_stripes_golden(oracle):
    st = synth.make_frame(W * 4, H * 4, 0, stripes=True)
    hdr = F.make_frame_headers(W * 4, H * 4)
    ri = hdr.rawi_hdr.raw_info
    needed, coef = oracle.stripes_compute(st, ri.black_level, ri.white_level, ri.frame_size)
    assert [needed] + list(coef) == list(G["stripes_coef"])
    out = oracle.stripes_apply(st, ri.black_level, ri.white_level, needed, coef)
    crc = [int(out.astype(np.uint64).sum()),
           int((out.astype(np.uint64) * np.arange(out.size).reshape(out.shape) % 65521).sum())]
    assert crc == [int(x) for x in G["stripes_out_crc"]]


def test_pattern_noise_golden(oracle):
    img = synth.make_frame(W, H, 1)
    rng = np.random.default_rng(1)
    img = (img.astype(np.int32) + rng.integers(-12, 13, size=(1, W)) + rng.integers(-9, 10, size=(H, 1)))
    img = img.clip(0, 16383).astype(np.uint16)
    img[H // 8:H // 4, W // 6:W // 3] = 15200
    assert np.array_equal(oracle.fix_pattern_noise(img, 15000), G["pn_out"])


def test_lj92_golden(oracle):
    stream = G["lj92_stream"]
    payload = np.concatenate([np.array([W * H * 2], dtype="<u4").view(np.uint8), stream])
    assert np.array_equal(oracle.lj92_decode_payload(payload, W, H), synth.make_frame(W, H, 2))


def test_chain_golden(oracle):
    w, h = W * 2, H * 2
    hdr = F.make_frame_headers(w, h)
    ri = hdr.rawi_hdr.raw_info
    frames = [synth.make_frame(w, h, i, hot_cold=True, stripes=True, bad_density=1e-3) for i in range(2)]
    got, _ = oracle.single_iso_chain(frames, ri.black_level, ri.white_level, ri.frame_size,
                                     chroma_smooth_method=3, fix_bad_pixels=1, fix_stripes=1)
    assert np.array_equal(np.stack(got), G["chain_out"])


# ---- dual ISO / AMaZE fixtures (tests/golden/dual_iso.npz, same generator script) ----------------------
GD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dual_iso.npz"))


def test_amaze_planes_golden(oracle):
    """AMaZE red/green/blue float planes, bit for bit (amaze_demosaic_RT.c:113-1487, SSE2 build)."""
    got = np.stack(oracle.amaze_demosaic(synth.amaze_test_mosaic(160, 104, 7)))
    assert np.array_equal(got.view(np.uint32), GD["amaze_rgb"].view(np.uint32))


@pytest.mark.parametrize("name,interp,cs,badpix", [("mean23", 1, 3, 1), ("amaze", 0, 0, 1)])
def test_dual_iso_golden(oracle, name, interp, cs, badpix):
    w, h = 256, 136
    img = synth.make_frame(w, h, 0, dual_iso=True, hot_cold=True, bad_density=1e-4)
    rc, got, _ = oracle.cr2hdr20(img, 2048, 15000, interp_method=interp, chroma_smooth_method=cs, fix_bad_pixels_mode=badpix)
    assert rc == 1
    assert np.array_equal(got, GD[f"diso_{name}_out"])
