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


# ---- LJ92 test-input generator (vectorised; fixed Huffman table) ---------------------------------
# Produces a valid lossless-JPEG stream (SOF3, one component, predictor 6) for the quadrant-interleaved
# frame, i.e. what a compressed MLV stores (reference main.c:617-681 decodes it with lj92.c).  Input
# generation only: it never touches the decode path.

# code length per SSSS category 0..16 (canonical code, Kraft sum < 1 so no all-ones code word)
_LJ_LEN = np.array([5, 4, 3, 3, 3, 3, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13], dtype=np.int64)


def _lj_table():
    order = sorted(range(17), key=lambda s: (_LJ_LEN[s], s))
    codes = np.zeros(17, dtype=np.int64)
    code, prev = 0, _LJ_LEN[order[0]]
    for s in order:
        code <<= int(_LJ_LEN[s] - prev)
        prev = _LJ_LEN[s]
        codes[s] = code
        code += 1
    counts = [int((_LJ_LEN == l).sum()) for l in range(1, 17)]
    return codes, counts, order


def lj92_encode_tiled(tiled, depth=14, predictor=6):
    """Encode an (already interleaved) uint16 [h, w] image; returns the JPEG stream as uint8."""
    t = np.ascontiguousarray(tiled, dtype=np.int64)
    h, w = t.shape
    pred = np.empty_like(t)
    pred[0, 0] = 1 << (depth - 1)
    pred[0, 1:] = t[0, :-1]
    pred[1:, 0] = t[:-1, 0]
    a, b, c = t[1:, :-1], t[:-1, 1:], t[:-1, :-1]                           # left, above, above-left
    pred[1:, 1:] = {1: a, 2: b, 3: c, 4: a + b - c, 5: a + ((b - c) >> 1), 6: b + ((a - c) >> 1),
                    7: (a + b) >> 1}[predictor]
    diff = (t - pred).reshape(-1)
    mag = np.abs(diff)
    ssss = np.zeros(diff.shape, dtype=np.int64)
    nz = mag > 0
    ssss[nz] = np.floor(np.log2(mag[nz])).astype(np.int64) + 1
    extra = np.where(diff < 0, diff + (1 << ssss) - 1, diff)
    codes, counts, order = _lj_table()
    nbits = _LJ_LEN[ssss] + ssss
    sym = (codes[ssss] << ssss) | extra                                     # code word followed by the magnitude bits
    ends = np.cumsum(nbits)
    total = int(ends[-1])
    starts = ends - nbits
    # expand every symbol into its bits (MSB first)
    idx = np.repeat(np.arange(sym.size), nbits)
    bitpos = np.arange(total) - np.repeat(starts, nbits)
    bits = ((sym[idx] >> (np.repeat(nbits, nbits) - 1 - bitpos)) & 1).astype(np.uint8)
    pad = (-total) % 8
    if pad:
        bits = np.concatenate([bits, np.ones(pad, np.uint8)])
    body = np.packbits(bits)
    ff = np.flatnonzero(body == 0xFF)
    if ff.size:                                                             # byte stuffing: 0xFF -> 0xFF 0x00
        body = np.insert(body, ff + 1, 0)
    head = bytearray([0xFF, 0xD8, 0xFF, 0xC3, 0, 11, depth, h >> 8, h & 255, w >> 8, w & 255, 1, 0, 0x11, 0,
                      0xFF, 0xC4, 0, 19 + 17, 0] + counts + order + [0xFF, 0xDA, 0, 8, 1, 0, 0, predictor, 0, 0])
    return np.concatenate([np.frombuffer(bytes(head), np.uint8), body, np.array([0xFF, 0xD9], np.uint8)])


def lj92_payload(img, depth=14, predictor=6):
    """VIDF payload of an LJ92 MLV frame: uint32 decoded size + stream of the interleaved frame."""
    h, w = img.shape
    stream = lj92_encode_tiled(quadrant_interleave(img), depth, predictor)
    return np.concatenate([np.array([w * h * 2], dtype="<u4").view(np.uint8), stream])


def amaze_test_mosaic(w, h, seed):
    """20-bit-range float mosaic with edges, texture at the Nyquist frequency and noise (AMaZE input as
    hdr.c:977-1026 prepares it: greens halved around black)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 30000 + 20000 * np.sin(xx / 7.0) * np.cos(yy / 5.0) + 100000 * ((xx // 32 + yy // 32) % 2) + rng.integers(-3000, 3000, (h, w))
    base = np.where((xx % 2) != (yy % 2), base * 0.5 + 60000, base)
    base[:, w // 2:] += 40000 * (xx[:, w // 2:] % 2)
    return np.clip(base, 0, 0xFFFFF).astype(np.int32).astype(np.float32)


def lzma_payload(img, bpp=14, *, lc=3, lp=0, pb=2, dict_size=1 << 22):
    """VIDF payload of a legacy LZMA clip (reference main.c:598-616): uint32 unpacked size, the 5 LZMA property
    bytes, then the raw LZMA1 stream of the packed frame.  Input generation only (Python's lzma module)."""
    import lzma
    packed = pack_bits(img, bpp).tobytes()
    alone = lzma.compress(packed, format=lzma.FORMAT_ALONE,
                          filters=[{"id": lzma.FILTER_LZMA1, "lc": lc, "lp": lp, "pb": pb, "dict_size": dict_size}])
    # FORMAT_ALONE = 5 property bytes + uint64 size + stream
    return np.frombuffer(np.uint32(len(packed)).tobytes() + alone[:5] + alone[13:], dtype=np.uint8).copy()
