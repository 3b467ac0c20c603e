"""GPU parity: LJ92-compressed VIDF payloads (BASELINE config 5 shape) vs the oracle decoder.  Bit-exact."""
import ctypes as C

import numpy as np
import pytest

import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F, synth

pytestmark = pytest.mark.gpu

LJ92 = F.VIDEO_CLASS_RAW | F.VIDEO_CLASS_FLAG_LJ92


@pytest.mark.parametrize("w,h", [(640, 360), (352, 98), (64, 34), (72, 38), (1920, 1080)])
def test_lj92_frame_matches_oracle(fresh_ctx, oracle, w, h):
    hdr = F.make_frame_headers(w, h, video_class=LJ92)
    img = synth.make_frame(w, h, 2, hot_cold=True, bad_density=1e-3)
    payload = oracle.lj92_payload(img)
    want = oracle.lj92_decode_payload(payload, w, h)
    assert np.array_equal(want, img)
    out, res = fresh_ctx.process_frame(hdr, payload, M.Options(), "lj92.MLV")
    assert res.status == 0
    assert np.array_equal(out, want)


@pytest.mark.parametrize("predictor,w,h", [(1, 640, 360), (2, 96, 70), (3, 96, 70), (4, 352, 98), (5, 96, 70), (7, 640, 360),
                                           (6, 70, 38)])
def test_lj92_wavefront_predictors(fresh_ctx, oracle, predictor, w, h):
    """Predictors other than 6 (and rasters the separated form does not take: width % 4 != 0) go through the
    skewed-wavefront kernel; 640x360 spans two blocks of 8 strips, so strips hand over across blocks."""
    hdr = F.make_frame_headers(w, h, video_class=LJ92)
    img = synth.make_frame(w, h, predictor, hot_cold=True, bad_density=1e-3)
    payload = synth.lj92_payload(img, predictor=predictor)
    want = oracle.lj92_decode_payload(payload, w, h)
    assert np.array_equal(want, img)
    out, res = fresh_ctx.process_frame(hdr, payload, M.Options(), "lj92pred.MLV")
    assert res.status == 0 and np.array_equal(out, want)


def test_lj92_reference_encoder_stream(fresh_ctx, oracle, ref):
    """A stream produced by the reference's own encoder (lj92.c:1104) decodes identically."""
    w, h = 480, 270
    hdr = F.make_frame_headers(w, h, video_class=LJ92)
    img = synth.make_frame(w, h, 7)
    tiled = np.ascontiguousarray(synth.quadrant_interleave(img))
    enc, n = C.POINTER(C.c_uint8)(), C.c_int()
    ref.lj92_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p]
    assert ref.lj92_encode(tiled.ctypes.data_as(C.c_void_p), w, h, 14, w * h, 0, None, 0, C.byref(enc), C.byref(n)) == 0
    stream = np.ctypeslib.as_array(enc, (n.value,)).copy()
    payload = np.concatenate([np.array([w * h * 2], dtype="<u4").view(np.uint8), stream])
    out, res = fresh_ctx.process_frame(hdr, payload, M.Options(), "lj92ref.MLV")
    assert np.array_equal(out, img)


def test_lj92_noisy_frame_long_codes_and_ff_stuffing(fresh_ctx, oracle):
    """Full-range noise forces long Huffman codes (> LUT width) and many 0xFF bytes in the stream."""
    w, h = 256, 64
    hdr = F.make_frame_headers(w, h, video_class=LJ92)
    rng = np.random.default_rng(5)
    img = rng.integers(0, 16384, size=(h, w), dtype=np.uint16)
    img[:, ::7] = rng.integers(8000, 8004, size=img[:, ::7].shape)
    payload = oracle.lj92_payload(img)
    assert (payload == 0xFF).sum() > 50
    out, res = fresh_ctx.process_frame(hdr, payload, M.Options(), "lj92noise.MLV")
    assert np.array_equal(out, img)


def test_lj92_then_corrections_and_batch(fresh_ctx, oracle):
    torch = pytest.importorskip("torch")
    w, h, n = 640, 360, 5
    hdr = F.make_frame_headers(w, h, video_class=LJ92)
    ri = hdr.rawi_hdr.raw_info
    frames = [synth.make_frame(w, h, i, hot_cold=True, stripes=True, bad_density=1e-4) for i in range(n)]
    want, _ = oracle.single_iso_chain(frames, ri.black_level, ri.white_level, ri.frame_size,
                                      chroma_smooth_method=3, fix_bad_pixels=1, fix_stripes=1)
    payloads = [oracle.lj92_payload(f) for f in frames]
    o = M.Options(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=1)
    for i in range(n):
        out, res = fresh_ctx.process_frame(hdr, payloads[i], o, "lj92chain.MLV")
        assert np.array_equal(out, want[i]), i
    # device batch with a common stride
    stride = (max(p.size for p in payloads) + 1024 + 15) // 16 * 16
    packed = np.zeros((n, stride), np.uint8)
    for i, p in enumerate(payloads):
        packed[i, :p.size] = p
    d_in = torch.from_numpy(packed).cuda()
    d_out = torch.empty((n, h * w), dtype=torch.int16, device="cuda")
    fresh_ctx.process_batch_device(hdr, o, "lj92chain.MLV", d_in.data_ptr(), stride, stride, d_out.data_ptr(), h * w, n,
                                   torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.uint16).reshape(n, h, w)
    for i in range(n):
        assert np.array_equal(got[i], want[i]), i


def test_lj92_fixed_length_symbols_still_converge(fresh_ctx, oracle):
    """Rows of identical +-2 steps: almost every symbol is 5 bits long, so wrong starts re-synchronise late
    or never; the boundary resolver must still reach the true parse (same case as tests/test_lj92_emu.py)."""
    w, h = 640, 360
    hdr = F.make_frame_headers(w, h, video_class=LJ92)
    row = 8192 + np.cumsum(np.where(np.arange(w) % 2 == 0, 2, -2))
    tiled = np.tile(row.astype(np.uint16), (h, 1))
    tiled[::37, ::53] += 3
    payload = np.concatenate([np.array([w * h * 2], dtype="<u4").view(np.uint8), synth.lj92_encode_tiled(tiled)])
    want = oracle.lj92_decode_payload(payload, w, h)
    out, res = fresh_ctx.process_frame(hdr, payload, M.Options(), "lj92fixed.MLV")
    assert res.status == 0 and np.array_equal(out, want)


def test_lj92_truncated_stream_fails_cleanly(fresh_ctx, oracle):
    w, h = 128, 64
    hdr = F.make_frame_headers(w, h, video_class=LJ92)
    payload = oracle.lj92_payload(synth.make_frame(w, h, 0))
    with pytest.raises(RuntimeError):
        fresh_ctx.process_frame(hdr, payload[: payload.size // 2].copy(), M.Options(), "cut.MLV")


def test_lj92_corrupt_stream_fails_cleanly(fresh_ctx, oracle):
    w, h = 128, 64
    hdr = F.make_frame_headers(w, h, video_class=LJ92)
    payload = oracle.lj92_payload(synth.make_frame(w, h, 0)).copy()
    payload[4:8] = 0            # destroy SOI
    with pytest.raises(RuntimeError):
        fresh_ctx.process_frame(hdr, payload, M.Options(), "bad.MLV")
