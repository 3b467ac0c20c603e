"""The parallel LJ92 decode programs (mlvfs_b200/csrc/lj92_core.cuh) executed on the host by
tests/emu/lj92_emu.cpp (one std::thread per CUDA thread, barriers for __syncthreads/__syncwarp, an
exchange array for shuffles) against the oracle decoder, bit for bit.  This pins the cooperative logic --
unstuffing, subsequence synchronisation across threads and blocks, the boundary resolver, the prefix
sums, the skewed prediction wavefront with its strip pipeline -- without a GPU; tests/test_gpu_lj92.py
then confirms the same source on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mlvfs_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "liblj92_emu.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-shared", "-fPIC", "-pthread", "-o", so,
                           os.path.join(ROOT, "tests", "emu", "lj92_emu.cpp")])
    lib = C.CDLL(so)
    lib.lj92_emu_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]

    def run(payload, w, h, flags=0, warps=4, parts=2):
        payload = np.ascontiguousarray(payload, np.uint8)
        out = np.full((h, w), 0xDEAD, np.uint16)
        stats = np.zeros(3, np.uint32)
        rc = lib.lj92_emu_decode(payload.ctypes.data, payload.size, w, h, out.ctypes.data, flags, warps, parts, stats.ctypes.data)
        return rc, out, stats

    return run


@pytest.mark.parametrize("w,h,flags", [(64, 34, 0), (64, 34, 4), (352, 98, 4), (640, 360, 0), (640, 360, 7)])
def test_parallel_decode_matches_oracle(emu, oracle, w, h, flags):
    """640x360 spans 7 decode blocks; flags 1|2 leave every block boundary to resolve_body and run the
    blocks in reverse order; flag 4 predicts with the wavefront instead of the separated predictor-6 form."""
    img = synth.make_frame(w, h, 2, hot_cold=True, bad_density=1e-3)
    payload = oracle.lj92_payload(img)
    want = oracle.lj92_decode_payload(payload, w, h)
    assert np.array_equal(want, img)
    rc, got, stats = emu(payload, w, h, flags)
    assert rc == 0
    assert np.array_equal(got, want)
    if (w, h) == (640, 360):
        assert stats[0] >= 5
        if flags & 1:
            assert stats[1] >= 1          # boundaries were really left stale for the resolver


def test_noise_long_codes_and_ff_runs(emu, oracle):
    """Full-range noise: Huffman codes longer than the first-level table, 0xFF bytes (also in runs)."""
    w, h = 256, 64
    rng = np.random.default_rng(5)
    img = rng.integers(0, 16384, size=(h, w), dtype=np.uint16)
    img[:, ::7] = rng.integers(8000, 8004, size=img[:, ::7].shape)
    payload = oracle.lj92_payload(img)
    assert (payload == 0xFF).sum() > 50
    for flags in (0, 4):
        rc, got, _ = emu(payload, w, h, flags)
        assert rc == 0 and np.array_equal(got, img)


def test_synth_encoder_and_ragged_width(emu, oracle):
    """The numpy encoder's fixed table; a width that is not a multiple of 32 takes the scalar row path."""
    for w, h in [(96, 40), (72, 38)]:
        img = synth.make_frame(w, h, 3)
        payload = synth.lj92_payload(img)
        want = oracle.lj92_decode_payload(payload, w, h)
        for flags in (0, 4):
            rc, got, _ = emu(payload, w, h, flags, warps=2)
            assert rc == 0 and np.array_equal(got, want)


def test_truncated_and_corrupt_streams_report_status(emu, oracle):
    w, h = 128, 64
    payload = oracle.lj92_payload(synth.make_frame(w, h, 0)).copy()
    bad = payload.copy()
    bad[4:8] = 0
    assert emu(bad, w, h)[0] == -1                       # no SOI
    cut = payload[: payload.size // 2].copy()
    assert emu(cut, w, h)[0] == -2                       # stream ends before the image does


@pytest.mark.parametrize("flags", [0, 1])
def test_fixed_length_symbols_still_converge(emu, oracle, flags):
    """Rows of identical +-2 steps: almost every symbol is 5 bits long, so a wrong start re-synchronises
    late or never inside its subsequence and the fix-up passes have to carry the true parse along."""
    w, h = 640, 360
    row = 8192 + np.cumsum(np.where(np.arange(w) % 2 == 0, 2, -2))
    tiled = np.tile(row.astype(np.uint16), (h, 1))
    tiled[::37, ::53] += 3                                           # a few other categories
    payload = np.concatenate([np.array([w * h * 2], dtype="<u4").view(np.uint8), synth.lj92_encode_tiled(tiled)])
    want = oracle.lj92_decode_payload(payload, w, h)
    rc, got, stats = emu(payload, w, h, flags)
    assert rc == 0 and np.array_equal(got, want)
    assert stats[0] >= 3


@pytest.mark.parametrize("predictor", [1, 2, 3, 4, 5, 7])
def test_other_predictors_take_the_wavefront(emu, oracle, predictor):
    w, h = 96, 70
    img = synth.make_frame(w, h, predictor)
    payload = synth.lj92_payload(img, predictor=predictor)
    want = oracle.lj92_decode_payload(payload, w, h)
    assert np.array_equal(want, img)
    rc, got, _ = emu(payload, w, h, warps=2)
    assert rc == 0 and np.array_equal(got, want)
