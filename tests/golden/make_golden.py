#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref) -- run in the build container.

The reference ships no golden vectors for this path, so these fixtures are produced by executing the
compiled reference itself on small deterministic inputs (mlvfs_b200/synth.py).  They travel with the
repo so the oracle restatement can be checked where /root/reference and oracle/_ref do not exist.
Each fixture stores the INPUT parameters (not the input pixels: they are regenerated from the seed) and
the reference's OUTPUT.

    python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mlvfs_b200 import mlvformat as F, synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
W, H = 192, 96


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def noisy(w, h, seed):
    img = synth.make_frame(w, h, seed)
    rng = np.random.default_rng(seed)
    img = (img.astype(np.int32) + rng.integers(-12, 13, size=(1, w)) + rng.integers(-9, 10, size=(h, 1)))
    img = img.clip(0, 16383).astype(np.uint16)
    img[h // 8:h // 4, w // 6:w // 3] = 15200
    return img


def main():
    ref = O.load_ref()
    assert ref is not None, "oracle/_ref is not built"
    hdr = F.make_frame_headers(W, H)
    g = {}
    # unpack at four depths
    for bpp in (8, 10, 12, 14):
        rng = np.random.default_rng(100 + bpp)
        img = rng.integers(0, 1 << bpp, size=(H, W), dtype=np.uint16)
        words = np.concatenate([synth.pack_bits(img, bpp), np.zeros(2, np.uint16)])
        h2 = F.make_frame_headers(W, H, bpp=bpp)
        out = np.zeros(W * H, np.uint16)
        ref.dng_get_image_data(C.byref(h2), p(words), p(out), 0, out.nbytes)
        g[f"unpack{bpp}_words"] = words
        g[f"unpack{bpp}_out"] = out
    # chroma smoothing
    img = synth.make_frame(W, H, 5)
    for m in (2, 3, 5):
        out = img.copy()
        ref.chroma_smooth(C.byref(hdr), p(out), m)
        g[f"cs{m}_out"] = out
    # bad pixels (normal + aggressive), clustered defects
    bp = synth.make_frame(W, H, 0, hot_cold=True, bad_density=2e-3)
    for k in range(6):
        bp[20 + 2 * k, 40] = 16000
        bp[60, 100 + 2 * k] = 16000
    g["badpix_in"] = bp
    for aggr in (0, 1):
        h2 = F.make_frame_headers(W, H, file_guid=0x7000 + aggr)
        out = bp.copy()
        with O.quiet_stdout():
            ref.fix_bad_pixels(C.byref(h2), p(out), aggr, 0)
        g[f"badpix{aggr}_out"] = out
    # stripes (unseeded rand() == srand(1))
    st = synth.make_frame(W * 4, H * 4, 0, stripes=True)
    h3 = F.make_frame_headers(W * 4, H * 4)
    C.CDLL(None).srand(1)
    ref.stripes_new_correction.argtypes = [C.c_char_p]
    corr = ref.stripes_new_correction(b"/golden/stripes.MLV")

    class Corr(C.Structure):
        _fields_ = [("next", C.c_void_p), ("name", C.c_char_p), ("needed", C.c_int), ("coef", C.c_int * 8)]

    ref.stripes_compute_correction.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_size_t]
    ref.stripes_apply_correction.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_size_t]
    ref.stripes_compute_correction(C.byref(h3), corr, p(st), 0, st.size)
    c = Corr.from_address(corr)
    g["stripes_coef"] = np.array([c.needed] + list(c.coef), np.int32)
    out = st.copy()
    ref.stripes_apply_correction(C.byref(h3), corr, p(out), 0, out.size)
    g["stripes_out_crc"] = np.array([int(out.astype(np.uint64).sum()), int((out.astype(np.uint64) * np.arange(out.size).reshape(out.shape) % 65521).sum())], np.uint64)
    # pattern noise
    pn = noisy(W, H, 1)
    out = pn.copy()
    with O.quiet_stdout():
        ref.fix_pattern_noise(p(out), W, H, 15000, 0)
    g["pn_out"] = out
    # LJ92: stream from the reference encoder + its own decode
    til = np.ascontiguousarray(synth.quadrant_interleave(synth.make_frame(W, H, 2)))
    enc, n = C.POINTER(C.c_uint8)(), C.c_int()
    ref.lj92_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p]
    ref.lj92_encode(p(til), W, H, 14, W * H, 0, None, 0, C.byref(enc), C.byref(n))
    g["lj92_stream"] = np.ctypeslib.as_array(enc, (n.value,)).copy()
    # whole C2 chain through process_frame on an MLV file
    with tempfile.TemporaryDirectory() as d:
        h4, frames = synth.make_clip(os.path.join(d, "G.MLV"), W * 2, H * 2, 2, variant=dict(hot_cold=True, stripes=True, bad_density=1e-3))
        ref.ref_set_mlv_dir(d.encode())
        ref.ref_set_options(3, 1, 1, 0, 0, 0, 0, 0, 0)
        C.CDLL(None).srand(1)
        outs = []
        with O.quiet_stdout():
            for i in range(2):
                o = np.zeros((H * 2, W * 2), np.uint16)
                ref.ref_process_frame(b"/G.MLV/G_%06d.dng" % i, p(o), o.nbytes, None)
                outs.append(o)
        g["chain_out"] = np.stack(outs)
    np.savez_compressed(os.path.join(OUT, "single_iso.npz"), **g)
    print("wrote", os.path.join(OUT, "single_iso.npz"), os.path.getsize(os.path.join(OUT, "single_iso.npz")), "bytes")
    dual_iso_golden(ref)
    focus_pixel_golden(ref)


FOCUS_CAM, FOCUS_RAW = 0x80000326, (1808, 727)
FOCUS_CASES = [  # (w, h, pan_x, pan_y, frame seed): full-width crop-mode frame, and a panned window whose left / right
    (1808, 727, 0, 0, 11),      # neighbours in the map fall outside [0, w) on the top / bottom rows (wrapped indices)
    (1280, 180, 260, 288, 12),
]


def focus_pixel_golden(ref):
    """fix_focus_pixels with one of the reference's REAL maps (mlvfs/data/80000326_1808x727.fpm).  The map's entries
    are stored in the fixture (the GPU box has no /root/reference); outputs are stored as (index, value) of the
    pixels the reference changed."""
    src = "/root/reference/mlvfs/data/%x_%dx%d.fpm" % (FOCUS_CAM, *FOCUS_RAW)
    xy = np.loadtxt(src, dtype=np.int32).reshape(-1, 2)
    g = {"fpm_xy": xy.astype(np.int16)}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            write_fpm(xy, FOCUS_CAM, *FOCUS_RAW)
            for k, (w, h, px, py, seed) in enumerate(FOCUS_CASES):
                hdr = F.make_frame_headers(w, h, camera_model=FOCUS_CAM, raw_width=FOCUS_RAW[0], raw_height=FOCUS_RAW[1],
                                           pan_x=px, pan_y=py)
                img = synth.make_frame(w, h, seed)
                for dual in (0, 1):
                    out = img.copy()
                    with O.quiet_stdout():
                        ref.fix_focus_pixels(C.byref(hdr), p(out), dual)
                    idx = np.flatnonzero(out != img).astype(np.int32)
                    g[f"case{k}_dual{dual}_idx"] = idx
                    g[f"case{k}_dual{dual}_val"] = out.reshape(-1)[idx]
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "focus_pixels.npz"), **g)
    print("wrote", os.path.join(OUT, "focus_pixels.npz"), os.path.getsize(os.path.join(OUT, "focus_pixels.npz")), "bytes")


def write_fpm(xy, cam, rw, rh, dirname="."):
    with open(os.path.join(dirname, "%x_%dx%d.fpm" % (cam, rw, rh)), "w") as f:
        for x, y in xy:
            f.write(f"{int(x)} \t {int(y)}\n")


def dual_iso_golden(ref):
    """AMaZE planes (amaze_demosaic_RT) and whole cr2hdr20_convert_data frames, mean23 and AMaZE-edge."""
    g = {}
    r, gr, b = O.ref_amaze_demosaic(synth.amaze_test_mosaic(160, 104, 7))
    g["amaze_rgb"] = np.stack([r, gr, b])
    ref.cr2hdr20_convert_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    w, h = 256, 136
    img = synth.make_frame(w, h, 0, dual_iso=True, hot_cold=True, bad_density=1e-4)
    for name, interp, cs, badpix in (("mean23", 1, 3, 1), ("amaze", 0, 0, 1)):
        hdr = F.make_frame_headers(w, h, file_guid=0x7100 + interp)
        out = img.copy()
        with O.quiet_stdout():
            rc = ref.cr2hdr20_convert_data(C.byref(hdr), p(out), interp, 1, 1, cs, badpix)
        assert rc == 1
        g[f"diso_{name}_out"] = out
    np.savez_compressed(os.path.join(OUT, "dual_iso.npz"), **g)
    print("wrote", os.path.join(OUT, "dual_iso.npz"), os.path.getsize(os.path.join(OUT, "dual_iso.npz")), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "focus":
        focus_pixel_golden(O.load_ref())
    else:
        main()
