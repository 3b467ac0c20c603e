"""Deterministic synthetic MLV input.

No sample footage ships with the reference and there is no network, so every test and bench input
is generated here from counter-based integer hashes (identical on every host, independent of the
numpy RNG implementation).  The generator only produces INPUT; expected output always comes from
the oracle.

Pixel model: scene s(x,y) = 40 + 1500*x/w + 600*checker64(x,y) + 300*y/h, RGGB colour gains
R 0.5 / G 1.0 / B 0.6, uniform noise +-16 DN, value = black + s + noise, soft-clipped at `white`.
Variants: hot/cold pixels (fixed per clip), 8-periodic column gains (vertical stripes), dual-ISO row
pairs (rows y%4 in {2,3} amplified 8x = 3 EV), focus-pixel geometry.
"""
import ctypes as C
import os

import numpy as np

from . import mlvformat as F

STRIPE_GAINS = np.array([1.0, 1.0, 1.012, 0.991, 1.008, 0.987, 1.015, 0.993])


def _hash32(x):
    """Avalanching 32-bit integer hash (vectorised, wraps mod 2^32)."""
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def _pixel_hash(w, h, salt):
    idx = np.arange(w * h, dtype=np.uint32).reshape(h, w)
    with np.errstate(over="ignore"):
        return _hash32(idx * np.uint32(0x9E3779B1) + np.uint32((salt * 0x85EBCA6B + 0x1234567) & 0xFFFFFFFF))


def make_frame(w, h, frame=0, *, black=2048, white=15000, bpp=14, hot_cold=False, stripes=False,
               dual_iso=False, bad_density=5e-6, noise_amp=16, seed=12345):
    """Return one synthetic Bayer frame as uint16 [h, w] (values < 2**bpp)."""
    yy, xx = np.mgrid[0:h, 0:w]
    checker = (((xx // 64) + (yy // 64)) & 1).astype(np.float64)
    s = 40.0 + 1500.0 * xx / w + 600.0 * checker + 300.0 * yy / h
    gain = np.where((yy & 1) == 0, np.where((xx & 1) == 0, 0.5, 1.0), np.where((xx & 1) == 0, 1.0, 0.6))
    s = s * gain
    if dual_iso:
        s = np.where((yy % 4) >= 2, s * 8.0, s)
    if stripes:
        s = s * STRIPE_GAINS[xx % 8]
    hsh = _pixel_hash(w, h, seed + frame)
    noise = (hsh % np.uint32(2 * noise_amp + 1)).astype(np.int64) - noise_amp
    v = black + np.floor(s).astype(np.int64) + noise
    jitter = ((hsh >> np.uint32(8)) % np.uint32(41)).astype(np.int64)
    v = np.where(v >= white, white + jitter, v)
    if hot_cold:
        # defect positions are a property of the sensor: fixed for the clip (salt independent of frame)
        dh = _pixel_hash(w, h, seed ^ 0x5BD1E995)
        thr = np.uint32(min(0xFFFFFFFF, int(bad_density * 2 ** 32)))
        hot = dh < thr
        cold = (dh >= thr) & (dh < np.uint32(min(0xFFFFFFFF, 2 * int(thr))))
        v = np.where(hot, 16000, v)
        v = np.where(cold, 1900, v)
    v = np.clip(v, 0, (1 << bpp) - 1)
    return v.astype(np.uint16)


def pack_bits(img, bpp=14):
    """Pack pixels MSB-first into the 16-bit little-endian word stream of raw.h:41-79.

    Returns uint16 words; the total bit count must be a multiple of 16.
    """
    px = np.ascontiguousarray(img, dtype=np.uint16).reshape(-1)
    n = px.size
    assert (n * bpp) % 16 == 0, "bit stream must end on a 16-bit word"
    if bpp == 14 and n % 8 == 0:
        p = px.reshape(-1, 8).astype(np.uint32)
        wds = np.empty((p.shape[0], 7), dtype=np.uint32)
        wds[:, 0] = (p[:, 0] << 2) | (p[:, 1] >> 12)
        wds[:, 1] = (p[:, 1] << 4) | (p[:, 2] >> 10)
        wds[:, 2] = (p[:, 2] << 6) | (p[:, 3] >> 8)
        wds[:, 3] = (p[:, 3] << 8) | (p[:, 4] >> 6)
        wds[:, 4] = (p[:, 4] << 10) | (p[:, 5] >> 4)
        wds[:, 5] = (p[:, 5] << 12) | (p[:, 6] >> 2)
        wds[:, 6] = (p[:, 6] << 14) | p[:, 7]
        return (wds & 0xFFFF).astype("<u2").reshape(-1)
    shifts = np.arange(bpp - 1, -1, -1, dtype=np.uint16)
    bits = ((px[:, None] >> shifts) & 1).astype(np.uint8).reshape(-1)
    be = np.packbits(bits)  # big-endian byte stream
    return be.view(">u2").astype("<u2")


def unpack_bits_numpy(words, n, bpp=14):
    """Independent numpy statement of the unpack (used to self-check pack_bits)."""
    be = np.ascontiguousarray(words, dtype="<u2").astype(">u2").view(np.uint8)
    bits = np.unpackbits(be)[: n * bpp].reshape(n, bpp).astype(np.uint32)
    weights = (1 << np.arange(bpp - 1, -1, -1)).astype(np.uint32)
    return (bits * weights).sum(axis=1).astype(np.uint16)


def quadrant_interleave(img):
    """Inverse of the LJ92 de-interleave of main.c:656-668: even rows first, even columns first."""
    return np.concatenate([np.concatenate([img[0::2, 0::2], img[0::2, 1::2]], axis=1),
                           np.concatenate([img[1::2, 0::2], img[1::2, 1::2]], axis=1)], axis=0)


def _blk(struct):
    return bytes(struct)


def write_mlv(path, frames_payload, headers, *, frame_space=0):
    """Write a single-chunk MLV: MLVI, RAWI, IDNT, then one VIDF block per payload.

    frames_payload: iterable of bytes-like VIDF payloads (packed bits, or u32 size + LJ92 stream).
    headers: FrameHeaders from mlvformat.make_frame_headers (file/rawi/idnt parts are written).
    """
    idx_path = os.path.splitext(path)[0] + ".IDX"
    if os.path.exists(idx_path):
        os.remove(idx_path)
    payloads = list(frames_payload)
    fh = F.clone_headers(headers)
    fh.file_hdr.videoFrameCount = len(payloads)
    with open(path, "wb") as f:
        f.write(_blk(fh.file_hdr))
        f.write(_blk(fh.rawi_hdr))
        f.write(_blk(fh.idnt_hdr))
        for i, pl in enumerate(payloads):
            pl = bytes(pl) if not isinstance(pl, (bytes, bytearray, memoryview)) else pl
            v = F.VidfHdr()
            C.memmove(C.byref(v), C.byref(fh.vidf_hdr), C.sizeof(F.VidfHdr))
            v.frameNumber = i
            v.timestamp = 1000 + i * 41666
            v.frameSpace = frame_space
            v.blockSize = C.sizeof(F.VidfHdr) + frame_space + len(pl)
            f.write(_blk(v))
            if frame_space:
                f.write(b"\0" * frame_space)
            f.write(pl)
    return path


def make_clip(path, w, h, nframes, *, headers=None, variant=None, bpp=14):
    """Generate frames + write an uncompressed MLV.  Returns (headers, [uint16 frames])."""
    variant = variant or {}
    if headers is None:
        headers = F.make_frame_headers(w, h, bpp=bpp)
    ri = headers.rawi_hdr.raw_info
    frames = [make_frame(w, h, i, black=ri.black_level, white=ri.white_level, bpp=bpp, **variant)
              for i in range(nframes)]
    write_mlv(path, (pack_bits(fr, bpp).tobytes() for fr in frames), headers)
    return headers, frames
