"""Oracle restatement of fix_focus_pixels (cs.c:440-503) vs tests/golden/focus_pixels.npz: the reference run on one of
its own real maps (mlvfs/data/80000326_1808x727.fpm), full frame and a panned crop whose out-of-frame neighbours act
on wrapped linear indices.  Runs anywhere."""
import os

import numpy as np
import pytest

from mlvfs_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "focus_pixels.npz"))
CAM, RAW = 0x80000326, (1808, 727)
CASES = [(1808, 727, 0, 0, 11), (1280, 180, 260, 288, 12)]       # tests/golden/make_golden.py FOCUS_CASES


def golden_frame(k, dual):
    w, h, px, py, seed = CASES[k]
    img = synth.make_frame(w, h, seed)
    want = img.copy()
    want.reshape(-1)[G[f"case{k}_dual{dual}_idx"]] = G[f"case{k}_dual{dual}_val"]
    return img, want, ((px + 7) & ~7, py & ~1)


@pytest.mark.parametrize("k", [0, 1])
@pytest.mark.parametrize("dual", [0, 1])
def test_oracle_focus_pixels_real_map(oracle, k, dual):
    img, want, crop = golden_frame(k, dual)
    got = oracle.focuspix_apply(img, 2048, G["fpm_xy"].astype(np.int32), crop=crop, dual_iso=dual)
    assert np.array_equal(got, want), int(np.count_nonzero(got != want))
