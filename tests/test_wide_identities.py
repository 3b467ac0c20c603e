"""Integer identities the wide fused kernel (mlvfs_b200/csrc/fused_wide.cuh) relies on, checked exhaustively on the
CPU against the oracle's restatement of stripes_apply_correction (stripes.c:250-266) and the plain definitions.

The kernel itself only runs on a GPU (tests/test_gpu_single_iso.py -k wide); these tests pin the arithmetic it was
rewritten into, so that a change of the host-side eligibility checks in fused.cu cannot silently admit inputs for which
the rewritten forms stop being exact.
"""
import numpy as np
import pytest

from oracle import pyoracle as O

EV = 32768


def _gain_high_half(v, coef, black, white):
    """FW_GAIN_X: the corrected sample is (v * gain + ((black << 16) - black * gain)) mod 2^32, clamped at
    white << 16 | 0xFFFF, read from the high half; samples <= black + 64 pass through."""
    v = v.astype(np.uint64)
    kx = ((black << 16) - black * coef) & 0xFFFFFFFF
    x = (v * np.uint64(coef) + np.uint64(kx)) & np.uint64(0xFFFFFFFF)
    x = np.minimum(x, np.uint64((white << 16) | 0xFFFF))
    return np.where(v > black + 64, x >> np.uint64(16), v).astype(np.uint16)


def _eligible(coef, black):
    """The host check in try_fused_single_iso (fused.cu): products fit 32 bits in both forms."""
    return 0 < coef < (1 << 18) and (16383 - black) * coef + (black << 16) < (1 << 32)


@pytest.mark.parametrize("black,white", [(2048, 15000), (1024, 16000), (4000, 14000), (0, 16383), (2048, 2100)])
def test_gain_in_high_half_equals_stripes_apply(black, white):
    rng = np.random.default_rng(black + white)
    coefs = [65536, 65535, 65537, 64700, 66500, 60000, 72000, 1, (1 << 17) + 12345, (1 << 18) - 1]
    coefs += [int(c) for c in rng.integers(64000, 67200, size=12)]
    v = np.arange(16384, dtype=np.uint16)
    img = np.repeat(v[None, :], 1, axis=0).reshape(-1, 8)            # every value once in each of the 8 columns' rows
    checked = 0
    for coef in coefs:
        if not _eligible(coef, black):
            continue
        c8 = [65536, 65536] + [coef] * 6                             # stripes.c:236-237: gains 0 and 1 are 1.0
        want = O.stripes_apply(img, black, white, 1, c8)
        got = img.copy()
        for col in range(2, 8):
            got[:, col] = _gain_high_half(img[:, col], coef, black, white)
        got[:, :2] = np.minimum(img[:, :2], white) if white > black + 64 else want[:, :2]
        assert np.array_equal(got, want), (coef, int(np.count_nonzero(got != want)))
        checked += 1
    assert checked >= 10


def test_eligibility_bound_is_tight_enough():
    """Inside the bound the 32-bit sum cannot wrap; just outside it, it can."""
    black = 2048
    coef_max = ((1 << 32) - 1 - (black << 16)) // (16383 - black)
    assert _eligible(min(coef_max, (1 << 18) - 1), black)
    assert (16383 - black) * (coef_max + 1) + (black << 16) >= (1 << 32)


def test_median_of_column_medians_through_shared_pair():
    """FW_MIDPAIR: med3(a, b, c) = max(min(a, b), min(max(a, b), c)) -- also for the INT_MIN entries raw2ev holds at
    black (main.c:154-179) and for wrapped differences."""
    rng = np.random.default_rng(7)
    x = rng.integers(-(1 << 31), 1 << 31, size=(200000, 3), dtype=np.int64).astype(np.int32)
    x[:1000, 0] = np.iinfo(np.int32).min
    x[1000:2000, 1] = np.iinfo(np.int32).min
    x[2000:3000] = rng.integers(-3, 3, size=(1000, 3))
    a, b, c = x[:, 0], x[:, 1], x[:, 2]
    got = np.maximum(np.minimum(a, b), np.minimum(np.maximum(a, b), c))
    assert np.array_equal(got, np.sort(x, axis=1)[:, 1])
    # the three-input form of the other medians: a + b + c - min - max in wrap-around arithmetic
    s = (a.astype(np.int64) + b + c - x.min(axis=1) - x.max(axis=1)).astype(np.int64)
    s = ((s + (1 << 31)) % (1 << 32) - (1 << 31)).astype(np.int32)
    assert np.array_equal(s, got)


def test_ev2raw_octaves_are_shifts_of_the_top_octave():
    """The wide kernel keeps only ev2raw[13 EV .. 14 EV) in shared memory: ev2raw[e] == top[e mod EV] >> (13 - e / EV)
    for every e in [0, 14 EV) (checked again on the device table when a context is created)."""
    import ctypes as C
    lib = O.load_oracle()
    lib.orc_ev2raw.restype = C.POINTER(C.c_int)                    # orc_luts.c: pointer to entry e = 0 (main.c:183-196)
    tab = np.ctypeslib.as_array(lib.orc_ev2raw(), shape=(14 * EV,))
    top = tab[13 * EV:14 * EV]
    e = np.arange(14 * EV)
    assert np.array_equal(tab, top[e % EV] >> (13 - e // EV))


def test_run_partition_covers_every_row_once():
    """The wide kernel's work split (nseg == 0): warp gw owns strip gw % nstrips and one contiguous run of that
    strip's nframes x ph quad rows, cut at frame ends.  Same integer arithmetic as the kernel."""
    for nframes, nstrips, ph, nwarps in [(256, 8, 540, 2368), (5, 8, 540, 2368), (3, 5, 47, 2368), (3, 16, 80, 2368),
                                         (8, 24, 1620, 2368), (7, 3, 181, 132 * 16), (1, 8, 540, 2368)]:
        seen = np.zeros((nframes, nstrips, ph), dtype=np.uint8)
        most = 0
        for gw in range(nwarps):
            strip = gw % nstrips
            nw_s = (nwarps - strip + nstrips - 1) // nstrips
            col = nframes * ph
            per = (col + nw_s - 1) // nw_s
            cur = min((gw // nstrips) * per, col)
            end = min(cur + per, col)
            rows = 0
            while cur < end:
                f = cur // ph
                q0 = cur - f * ph
                q1 = min(q0 + (end - cur), ph)
                seen[f, strip, q0:q1] += 1
                rows += q1 - q0 + 2
                cur += q1 - q0
            most = max(most, rows)
        assert (seen == 1).all()
        assert most <= -(-nframes * nstrips * ph // nwarps) + 8 + 2 * (1 + per // ph)
