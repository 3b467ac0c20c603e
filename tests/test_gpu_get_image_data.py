"""GPU parity: the exported drop-in get_image_data(frame_headers, FILE *, out, offset, max_size) (reference
main.c:569-706, the symbol gif.c:164 and the Pismo front-end call): uncompressed clips incl. ranged reads
(offset != 0, what win/mlvfs-pfm.cpp:1221-1237 issues), LJ92 clips (decode + the quadrant de-interleave of
main.c:656-668) and legacy LZMA clips (main.c:598-616), and the same clips through mlvb_process_frame.
Bit-exact against the frames the clips were made from and, where the compiled reference is present, against its
own get_image_data on the same FILE."""
import ctypes as C
import os

import numpy as np
import pytest

import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F, synth

pytestmark = pytest.mark.gpu

libc = C.CDLL(None)
libc.fopen.restype = C.c_void_p
libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
libc.fclose.argtypes = [C.c_void_p]


def make_clip(path, w, h, n, codec, bpp=14):
    vc = F.VIDEO_CLASS_RAW | {"raw": 0, "lj92": F.VIDEO_CLASS_FLAG_LJ92, "lzma": F.VIDEO_CLASS_FLAG_LZMA}[codec]
    hdr = F.make_frame_headers(w, h, video_class=vc, bpp=bpp)
    frames = [synth.make_frame(w, h, i, hot_cold=True, bpp=bpp) for i in range(n)]
    enc = {"raw": lambda f: synth.pack_bits(f, bpp).tobytes(), "lj92": lambda f: synth.lj92_payload(f).tobytes(),
           "lzma": lambda f: synth.lzma_payload(f, bpp).tobytes()}[codec]
    payloads = [enc(f) for f in frames]
    synth.write_mlv(path, payloads, hdr, frame_space=8 if codec == "raw" else 0)
    # per-frame headers the way mlv_get_frame_headers fills them: file position of the VIDF block + its header
    hdrs, pos = [], C.sizeof(F.FileHdr) + C.sizeof(F.RawiHdr) + C.sizeof(F.IdntHdr)
    for i, pl in enumerate(payloads):
        fh = F.clone_headers(hdr)
        fh.position = pos
        fh.vidf_hdr.frameNumber = i
        fh.vidf_hdr.frameSpace = 8 if codec == "raw" else 0
        fh.vidf_hdr.blockSize = C.sizeof(F.VidfHdr) + fh.vidf_hdr.frameSpace + len(pl)
        pos += fh.vidf_hdr.blockSize
        hdrs.append(fh)
    return hdrs, frames


def call(lib, fh, fp, offset, max_size, cap):
    out = np.full(cap, 0xAB, np.uint8)
    lib.get_image_data.restype = C.c_size_t
    lib.get_image_data.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_size_t]
    n = lib.get_image_data(C.byref(fh), fp, out.ctypes.data_as(C.c_void_p), offset, max_size)
    return n, out


@pytest.mark.parametrize("codec,w,h", [("raw", 640, 360), ("lj92", 640, 360), ("lj92", 1920, 1080), ("lzma", 640, 360)])
def test_get_image_data_whole_frame(fresh_ctx, oracle, tmp_path, codec, w, h):
    path = str(tmp_path / "G.MLV")
    hdrs, frames = make_clip(path, w, h, 2, codec)
    ref = oracle.load_ref()
    fp = libc.fopen(path.encode(), b"rb")
    try:
        for fh, fr in zip(hdrs, frames):
            n, out = call(M.lib(), fh, fp, 0, w * h * 2, w * h * 2)
            assert n == w * h * 2
            assert np.array_equal(out.view(np.uint16).reshape(h, w), fr)
            if ref is not None:
                with oracle.quiet_stdout():
                    rn, rout = call(ref, fh, fp, 0, w * h * 2, w * h * 2)
                assert np.array_equal(rout, out)
                assert rn == (0 if codec == "lj92" else w * h * 2)      # main.c:617-681 never assigns `result` for LJ92
    finally:
        libc.fclose(fp)


@pytest.mark.parametrize("codec", ["raw", "lzma", "lj92"])
def test_get_image_data_ranged_reads(fresh_ctx, oracle, tmp_path, codec):
    w, h = 640, 360
    path = str(tmp_path / "R.MLV")
    hdrs, frames = make_clip(path, w, h, 1, codec)
    ref = oracle.load_ref() if codec == "raw" else None                 # the reference's ranged reads are only sound for raw
    flat = frames[0].reshape(-1).view(np.uint8)
    fp = libc.fopen(path.encode(), b"rb")
    try:
        for offset, size in [(0, 4096), (2 * 1000, 4096), (2 * 12345, 65536), (w * h * 2 - 1024, 1024), (-512, 4096)]:
            n, out = call(M.lib(), hdrs[0], fp, offset, size, size + 16)
            assert n == size
            skip = -offset if offset < 0 else 0
            first = max(offset, 0)
            assert np.array_equal(out[skip:size], flat[first:first + size - skip])
            assert np.all(out[size:] == 0xAB) and np.all(out[:skip] == 0xAB)
            if ref is not None:
                rn, rout = call(ref, hdrs[0], fp, offset, size, size + 16)
                assert rn == n and np.array_equal(rout, out)
    finally:
        libc.fclose(fp)


def test_get_image_data_corrupt_lj92_fails(fresh_ctx, tmp_path):
    w, h = 640, 360
    path = str(tmp_path / "C.MLV")
    hdrs, frames = make_clip(path, w, h, 1, "lj92")
    with open(path, "r+b") as f:                                        # break the SOF3 marker of the stream
        f.seek(hdrs[0].position + C.sizeof(F.VidfHdr) + 4 + 2)
        f.write(b"\x00\x00\x00\x00")
    fp = libc.fopen(path.encode(), b"rb")
    try:
        n, out = call(M.lib(), hdrs[0], fp, 0, w * h * 2, w * h * 2)
        assert n == 0
    finally:
        libc.fclose(fp)


@pytest.mark.parametrize("bpp", [14, 12])
def test_lzma_clip_through_process_frame(fresh_ctx, oracle, bpp):
    """mlvb_process_frame on an LZMA payload: expanded on the host, then the usual GPU chain (here cs2x2)."""
    w, h = 640, 360
    hdr = F.make_frame_headers(w, h, video_class=F.VIDEO_CLASS_RAW | F.VIDEO_CLASS_FLAG_LZMA, bpp=bpp)
    img = synth.make_frame(w, h, 7, bpp=bpp)
    black = hdr.rawi_hdr.raw_info.black_level
    out, res = fresh_ctx.process_frame(hdr, synth.lzma_payload(img, bpp), M.Options(chroma_smooth=2), "lzma.MLV")
    assert res.status == 0
    assert np.array_equal(out, oracle.chroma_smooth(img, black, 2))
    bad = synth.lzma_payload(img, bpp)
    bad[9 + 40:9 + 60] ^= 0x5A
    with pytest.raises(RuntimeError):
        fresh_ctx.process_frame(hdr, bad[:2000], M.Options(), "lzma.MLV")
