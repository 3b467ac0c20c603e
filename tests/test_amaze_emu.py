"""The CUDA AMaZE tile program (mlvfs_b200/csrc/amaze_tile.cuh) executed on the host by tests/emu/amaze_emu.cpp
(one std::thread per CUDA thread, std::barrier for __syncthreads) against the oracle, bit for bit.  This pins the
block-cooperative logic -- pass order, the row-sequential passes, over-run lanes, workspace reuse -- without a GPU;
tests/test_gpu_amaze.py then only has to confirm that the device arithmetic rounds the same way."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mlvfs_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libamaze_emu.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-o", so,
                           os.path.join(ROOT, "tests", "emu", "amaze_emu.cpp")])
    lib = C.CDLL(so)
    lib.amaze_emu.argtypes = [C.c_void_p] * 4 + [C.c_int] * 6

    def run(raw, nthr=256, order=0, poison=1):
        h, w = raw.shape
        ws = w + 16
        src = np.zeros((h, ws), np.float32)
        src[:, :w] = raw
        outs = [np.full((h, ws), np.nan, np.float32) for _ in range(3)]
        lib.amaze_emu(src.ctypes.data, outs[0].ctypes.data, outs[1].ctypes.data, outs[2].ctypes.data, ws, w, h, nthr, order, poison)
        return [o[:, :w].copy() for o in outs]

    return run


@pytest.mark.parametrize("w,h,order", [(160, 160, 2), (256, 206, 1)])
def test_tile_program_matches_oracle(emu, oracle, w, h, order):
    """(256, 206): partial bottom tile, the bottom-border overrun (h - top = 158) and a 32-wide no-output tile;
    thread ids are permuted (`order`) and the workspace is NaN-poisoned before every tile."""
    raw = synth.amaze_test_mosaic(w, h, w * 7 + h)
    want = oracle.amaze_demosaic(raw, fresh_tiles=1)
    got = emu(raw, order=order)
    for name, a, b in zip("RGB", got, want):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), name
