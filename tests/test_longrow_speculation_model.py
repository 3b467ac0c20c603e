"""The algorithm behind pixfix.cu's warp-wide walk of a dense bad-pixel row (walk_dense_warp), modelled in Python against
the serial list-order walk of the reference (cs.c:87-109 through the oracle's LUTs): 32 pieces walked speculatively after
a warm-up from the unrepaired pixels, verified in order against the true state, failed pieces re-walked until they rejoin
their earlier results.  The kernel source itself is checked on the GPU (tests/test_gpu_dual_iso.py::
test_dual_iso_dense_rows_are_walked_by_the_whole_warp); this test pins the argument: the scheme is exact whatever the
content, the speculation only decides how much is redone."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as O

BLACK, EVR = 2048, 32768
WARM, LANES = 128, 32


@pytest.fixture(scope="module")
def luts():
    lib = O.load_oracle()
    lib.orc_raw2ev.argtypes = [C.c_int]
    r2e = np.ctypeslib.as_array(lib.orc_raw2ev(BLACK), shape=(16384,)).astype(np.int64).copy()
    e2r = np.ctypeslib.as_array(lib.orc_ev2raw(), shape=(14 * EVR,)).astype(np.int64).copy()
    return r2e, e2r


def step(row, x, e0, e1, e2, luts):
    """One entry of interpolate_horizontal (cs.c:87-109) given the EVs of the three repaired pixels to the left."""
    r2e, e2r = luts
    o1, o2, o3 = r2e[row[x + 1]], r2e[row[x + 2]], r2e[row[x + 3]]
    d1, d2 = abs(o3 - o1), abs(e2 - e0)
    s = d1 + d2
    if s == 0:
        return int(row[x + 2])
    c1, c2 = ((s - d1) << 8) // s, ((s - d2) << 8) // s
    ev = ((o2 * c1) >> 8) + ((e1 * c2) >> 8)
    return int((e2r[min(max(ev, 0), 14 * EVR - 1)] + BLACK) & 0xFFFF)


def serial(row, x0, cnt, luts):
    r2e = luts[0]
    out = row.copy()
    for x in range(x0, x0 + cnt):
        out[x] = step(row, x, r2e[out[x - 3]], r2e[out[x - 2]], r2e[out[x - 1]], luts)   # originals to the right, results to the left
    return out


def speculative(row, x0, cnt, luts):
    """walk_dense_warp: returns (result row, pieces that failed verification, columns re-walked)."""
    r2e = luts[0]
    L = (cnt + LANES - 1) // LANES
    out = row.copy()
    a, b, spans = {}, {}, []
    for k in range(LANES):
        p0, p1 = x0 + k * L, min(x0 + (k + 1) * L, x0 + cnt)
        if p0 >= x0 + cnt:
            break
        xb = max(x0, p0 - WARM)
        st = [r2e[row[xb - 3]], r2e[row[xb - 2]], r2e[row[xb - 1]]]       # the guess: unrepaired pixels (true at x0)
        for x in range(xb, p1):
            if x == p0:
                a[k] = tuple(st)
            v = step(row, x, st[0], st[1], st[2], luts)
            if x >= p0:
                out[x] = v
            st = [st[1], st[2], r2e[v]]
        b[k] = tuple(st)
        spans.append((k, p0, p1, xb))
    failed = redone = 0
    for k, p0, p1, xb in spans:                                          # verification in order
        if xb == x0 or a[k] == b[k - 1]:
            continue
        failed += 1
        st, same = list(b[k - 1]), 0
        for x in range(p0, p1):
            v = step(row, x, st[0], st[1], st[2], luts)
            redone += 1
            same = same + 1 if v == out[x] else 0
            if same >= 3:
                break                                                    # back on the earlier trajectory: b[k] stands
            out[x] = v
            st = [st[1], st[2], r2e[v]]
        else:
            b[k] = tuple(st)
    return out, failed, redone


@pytest.mark.parametrize("content", ["frame", "steps", "noise", "flat", "ramp"])
def test_speculative_pieces_equal_the_serial_walk(luts, content):
    from mlvfs_b200 import synth
    w, x0 = 2600, 7
    cnt = w - 3 - x0 - 4
    rng = np.random.default_rng(3)
    xs = np.arange(w)
    if content == "frame":
        row = synth.make_frame(w, 8, 8, dual_iso=True)[6].astype(np.int64)
    elif content == "steps":
        row = 2200 + 9000 * ((xs // 2) % 2) + rng.integers(0, 3, w)
    elif content == "noise":
        row = rng.integers(1900, 16383, w)
    elif content == "flat":
        row = np.full(w, 5000)
        row[::97] = 9000
    else:
        row = 2100 + (xs * 5) % 14000 + rng.integers(0, 2, w)
    row = row.astype(np.int64)
    want = serial(row, x0, cnt, luts)
    got, failed, redone = speculative(row, x0, cnt, luts)
    assert np.array_equal(got, want), (content, int(np.count_nonzero(got != want)))
    print(f"{content}: {failed} of {LANES} pieces re-walked, {redone} columns redone of {cnt}")
    if content == "frame":
        assert redone < cnt // 2, "the speculation should converge on ordinary content"
