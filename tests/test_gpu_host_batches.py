"""GPU parity: mlvb_process_frames, the host-batch entry the --prefetch queue calls (include/mlvfs_b200.h): frames of
one clip in host memory -> one device batch (the wide fused kernel at 1080p) -> host destinations.  Bit-exact
against the oracle chain; also pageable buffers, LJ92 batches, concurrent batches from several threads, and option
sets that are pipelined frame by frame instead (dual ISO, deflicker)."""
import ctypes as C
import threading

import numpy as np
import pytest

import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F, synth

pytestmark = pytest.mark.gpu


def _headers(hdr, n):
    out = []
    for i in range(n):
        h = F.clone_headers(hdr)
        h.vidf_hdr.frameNumber = i
        out.append(h)
    return out


def test_host_batch_single_iso_chain_uses_the_wide_kernel(fresh_ctx, oracle):
    w, h, n = 1920, 1080, 9
    hdr = F.make_frame_headers(w, h, file_guid=0xBA7C)
    ri = hdr.rawi_hdr.raw_info
    frames = [synth.make_frame(w, h, i, hot_cold=True, stripes=True) for i in range(n)]
    o = M.Options(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=1)
    want, _ = oracle.single_iso_chain(frames, ri.black_level, ri.white_level, ri.frame_size, chroma_smooth_method=3,
                                      fix_bad_pixels=1, fix_stripes=1)
    packed = [synth.pack_bits(f) for f in frames]
    out0, res0 = fresh_ctx.process_frame(hdr, packed[0], o, "hb.MLV")           # frame 0 creates the per-clip state
    assert np.array_equal(out0, want[0])
    nbytes = packed[0].nbytes
    pin_in = M.PinnedBuffer(n * nbytes)
    pin_in.array[:] = np.concatenate([p.view(np.uint8) for p in packed])
    pin_out = M.PinnedBuffer(n * w * h * 2)
    try:
        wide0, hb0 = fresh_ctx.path_count(1), fresh_ctx.path_count(2)
        rc, res = fresh_ctx.process_frames(_headers(hdr, n), [pin_in.ptr + i * nbytes for i in range(n)], [nbytes] * n, o, "hb.MLV",
                                           [pin_out.ptr + i * w * h * 2 for i in range(n)])
        assert rc == 0 and all(r.status == 0 and r.black_level == 2048 for r in res)
        got = pin_out.array.view(np.uint16).reshape(n, h, w)
        for i in range(n):
            assert np.array_equal(got[i], want[i]), i
        assert fresh_ctx.path_count(2) == hb0 + 1 and fresh_ctx.path_count(1) == wide0 + 1
        # pageable sources and destinations, a batch of 2, and a batch of 1 (falls back to the frame pipeline)
        for k in (2, 1):
            outs = [np.zeros((h, w), np.uint16) for _ in range(k)]
            rc, res = fresh_ctx.process_frames(_headers(hdr, k), [packed[3 + i].ctypes.data for i in range(k)], [nbytes] * k, o, "hb.MLV",
                                               [x.ctypes.data for x in outs])
            assert rc == 0
            for i in range(k):
                assert np.array_equal(outs[i], want[3 + i])
    finally:
        pin_in.free()
        pin_out.free()


def test_host_batches_from_several_threads(fresh_ctx, oracle):
    w, h, n, T = 1280, 720, 6, 4
    hdr = F.make_frame_headers(w, h, file_guid=0xBA7D)
    frames = [synth.make_frame(w, h, i, hot_cold=True) for i in range(n * T)]
    o = M.Options(chroma_smooth=3, fix_bad_pixels=1)
    lst = oracle.badpix_detect(frames[0], 2048, 0)
    want = [oracle.chroma_smooth(oracle.badpix_apply(f, 2048, lst), 2048, 3) for f in frames]
    packed = [synth.pack_bits(f) for f in frames]
    fresh_ctx.process_frame(hdr, packed[0], o, "hbt.MLV")
    nbytes = packed[0].nbytes
    outs = [np.zeros((h, w), np.uint16) for _ in frames]
    errs = []

    def worker(t):
        for rep in range(3):
            idx = list(range(t * n, (t + 1) * n))
            rc, res = fresh_ctx.process_frames(_headers(hdr, n), [packed[i].ctypes.data for i in idx], [nbytes] * n, o, "hbt.MLV",
                                               [outs[i].ctypes.data for i in idx])
            if rc != 0:
                errs.append((t, rc))

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    for i in range(n * T):
        assert np.array_equal(outs[i], want[i]), i


def test_host_batch_lj92_and_corrupt_frame(fresh_ctx):
    w, h, n = 1280, 720, 5
    hdr = F.make_frame_headers(w, h, video_class=F.VIDEO_CLASS_RAW | F.VIDEO_CLASS_FLAG_LJ92)
    frames = [synth.make_frame(w, h, i) for i in range(n)]
    pls = [synth.lj92_payload(f) for f in frames]
    outs = [np.zeros((h, w), np.uint16) for _ in frames]
    hb0 = fresh_ctx.path_count(2)
    rc, res = fresh_ctx.process_frames(_headers(hdr, n), [p.ctypes.data for p in pls], [p.nbytes for p in pls], M.Options(), "hbl.MLV",
                                       [x.ctypes.data for x in outs])
    assert rc == 0 and fresh_ctx.path_count(2) == hb0 + 1
    for i in range(n):
        assert np.array_equal(outs[i], frames[i])
    bad = pls[2].copy()
    bad[6:10] = 0                                                               # break the SOF3 marker
    pls2 = pls[:2] + [bad] + pls[3:]
    rc, res = fresh_ctx.process_frames(_headers(hdr, n), [p.ctypes.data for p in pls2], [p.nbytes for p in pls2], M.Options(), "hbl.MLV",
                                       [x.ctypes.data for x in outs])
    assert rc != 0 and res[2].status != 0 and res[0].status == 0 and res[4].status == 0


def test_host_batch_dual_iso_and_deflicker_are_pipelined_per_frame(fresh_ctx, oracle):
    w, h, n = 640, 384, 6
    hdr = F.make_frame_headers(w, h, file_guid=0xBA7E)
    o = M.Options(dual_iso=2, hdr_interpolation_method=1, chroma_smooth=3, deflicker=3000)
    frames = [synth.make_frame(w, h, i, dual_iso=True) for i in range(n)]
    packed = [synth.pack_bits(f) for f in frames]
    st = oracle.new_diso_state()
    want = [oracle.cr2hdr20(f, 2048, 15000, interp_method=1, chroma_smooth_method=3, state=st)[1] for f in frames]
    outs = [np.zeros((h, w), np.uint16) for _ in frames]
    hb0 = fresh_ctx.path_count(2)
    rc, res = fresh_ctx.process_frames(_headers(hdr, n), [p.ctypes.data for p in packed], [p.nbytes for p in packed], o, "hbd.MLV",
                                       [x.ctypes.data for x in outs])
    assert rc == 0 and fresh_ctx.path_count(2) == hb0
    for i in range(n):
        assert res[i].is_dual_iso == 1 and res[i].black_level == 8192
        assert (res[i].exposure_bias[0], res[i].exposure_bias[1]) == oracle.deflicker(frames[i], 14, 2048, 3000)
        assert np.abs(outs[i].astype(np.int32) - want[i].astype(np.int32)).max() <= 1
