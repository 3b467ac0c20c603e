"""GPU parity: pattern-noise removal (patternnoise.c) vs the oracle.  Integer medians: bit-exact."""
import numpy as np
import pytest

import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F, synth

pytestmark = pytest.mark.gpu


def noisy_frame(w, h, seed=1):
    img = synth.make_frame(w, h, seed)
    rng = np.random.default_rng(seed)
    img = (img.astype(np.int32) + rng.integers(-12, 13, size=(1, w)) + rng.integers(-9, 10, size=(h, 1)))
    img = img.clip(0, 16383).astype(np.uint16)
    img[h // 8:h // 4, w // 6:w // 3] = 15200        # clipped highlight: masked out (>= white)
    return img


@pytest.mark.parametrize("w,h", [(640, 360), (256, 130), (1920, 1080)])
def test_fix_pattern_noise_dropin(fresh_ctx, oracle, w, h):
    img = noisy_frame(w, h)
    want = oracle.fix_pattern_noise(img, 15000)
    got = img.copy()
    M.lib().fix_pattern_noise(got.ctypes.data, w, h, 15000, 0)
    assert (want != img).sum() > w * h // 4
    assert np.array_equal(got, want)


def test_pattern_noise_in_pipeline_order(fresh_ctx, oracle):
    """--fix-pattern-noise runs right after the unpack, before bad-pixel / chroma / stripes (main.c:946-997)."""
    w, h = 640, 360
    hdr = F.make_frame_headers(w, h)
    ri = hdr.rawi_hdr.raw_info
    img = noisy_frame(w, h, 3)
    pn = oracle.fix_pattern_noise(img, ri.white_level)
    want, _ = oracle.single_iso_chain([pn], ri.black_level, ri.white_level, ri.frame_size, chroma_smooth_method=2,
                                      fix_bad_pixels=1, fix_stripes=0)
    o = M.Options(fix_pattern_noise=1, chroma_smooth=2, fix_bad_pixels=1)
    out, res = fresh_ctx.process_frame(hdr, synth.pack_bits(img), o, "pn.MLV")
    assert np.array_equal(out, want[0])
