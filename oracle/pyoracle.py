"""ctypes loaders for the CPU oracle -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this.  The product package (mlvfs_b200/) never does.

  load_oracle()  -> oracle/liboracle.so          (our C restatement, oracle/*.c)
  load_ref()     -> oracle/_ref/libmlvfs_ref.so  (the unmodified reference compiled by oracle/Makefile;
                                                  None when it was never built)
"""
import contextlib
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmlvfs_ref.so")
REFERENCE_SRC = "/root/reference/mlvfs"


class Pixel(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int)]


class RandState(C.Structure):
    _fields_ = [("r", C.c_uint32 * 34), ("f", C.c_int), ("b", C.c_int), ("primed", C.c_int)]


def build(ref=True, quiet=True):
    """Compile liboracle.so and, where the reference sources exist, oracle/_ref."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"], stdout=out)
    if ref and os.path.isdir(REFERENCE_SRC):
        subprocess.check_call(["make", "-s", "-C", HERE, "-j8", "ref"], stdout=out)


_oracle = None
_ref = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def load_oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        lib = C.CDLL(ORACLE_SO)
        lib.orc_raw2ev.restype = C.POINTER(C.c_int)
        lib.orc_raw2evf.restype = C.POINTER(C.c_double)
        lib.orc_ev2raw.restype = C.POINTER(C.c_int)
        lib.orc_unpack.restype = C.c_size_t
        lib.orc_unpack.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_size_t, C.c_int]
        lib.orc_badpix_detect.restype = C.c_size_t
        lib.orc_badpix_detect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_size_t]
        lib.orc_badpix_apply.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                         C.c_int, C.c_int, C.c_int]
        lib.orc_focuspix_apply.argtypes = lib.orc_badpix_apply.argtypes
        lib.orc_chroma_smooth_u16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.orc_chroma_smooth_u32.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_void_p]
        lib.orc_stripes_compute.restype = C.c_int
        lib.orc_stripes_compute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p]
        lib.orc_stripes_apply.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.orc_rand_seed.argtypes = [C.c_void_p, C.c_uint]
        lib.orc_rand_next.argtypes = [C.c_void_p]
        _oracle = lib
    return _oracle


def load_ref():
    """The compiled reference, or None if oracle/_ref was not built (no /root/reference at build time)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            if os.path.isdir(REFERENCE_SRC):
                build(ref=True)
            else:
                return None
        lib = C.CDLL(REF_SO)
        lib.ref_process_frame.restype = C.c_size_t
        lib.ref_process_frame.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.ref_sizeof_frame_headers.restype = C.c_size_t
        lib.ref_offsetof_frame_headers.restype = C.c_size_t
        lib.dng_get_image_data.restype = C.c_size_t
        lib.dng_get_image_data.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_size_t]
        lib.get_raw2ev.restype = C.POINTER(C.c_int)
        lib.get_ev2raw.restype = C.POINTER(C.c_int)
        lib.get_raw2evf.restype = C.POINTER(C.c_double)
        lib.stripes_new_correction.restype = C.c_void_p
        lib.stripes_get_correction.restype = C.c_void_p
        lib.ref_init()
        _ref = lib
    return _ref


@contextlib.contextmanager
def quiet_stdout():
    """The reference printf()s per frame; silence fd 1 around calls into it."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


# ---- numpy-friendly wrappers around the restatement -------------------------------------------

def unpack(words, npix, bpp=14, offset=0, nbytes=None):
    lib = load_oracle()
    words = np.ascontiguousarray(words, dtype=np.uint16)
    nbytes = npix * 2 if nbytes is None else nbytes
    out = np.zeros(nbytes // 2, dtype=np.uint16)
    first_word = (max(0, offset) // 2) * bpp // 16
    lib.orc_unpack(C.c_void_p(words.ctypes.data + 2 * first_word), _p(out), offset, nbytes, bpp)
    return out


def chroma_smooth(img, black, method):
    lib = load_oracle()
    out = np.ascontiguousarray(img, dtype=np.uint16).copy()
    h, w = out.shape
    lib.orc_chroma_smooth_u16(_p(out), w, h, black, method)
    return out


def badpix_detect(img, black, aggressive, crop=(0, 0)):
    lib = load_oracle()
    img = np.ascontiguousarray(img, dtype=np.uint16)
    h, w = img.shape
    cap = 1 << 16
    while True:
        arr = (Pixel * cap)()
        n = lib.orc_badpix_detect(_p(img), w, h, black, int(aggressive), crop[0], crop[1], arr, cap)
        if n <= cap:
            break
        cap = int(n)
    return np.frombuffer(arr, dtype=np.int32, count=2 * n).reshape(n, 2).copy()


def badpix_apply(img, black, plist, crop=(0, 0), dual_iso=0):
    lib = load_oracle()
    out = np.ascontiguousarray(img, dtype=np.uint16).copy()
    h, w = out.shape
    pl = np.ascontiguousarray(plist, dtype=np.int32)
    lib.orc_badpix_apply(_p(out), w, h, black, _p(pl), len(pl), crop[0], crop[1], dual_iso)
    return out


def focuspix_apply(img, black, pmap, crop=(0, 0), dual_iso=0):
    lib = load_oracle()
    out = np.ascontiguousarray(img, dtype=np.uint16).copy()
    h, w = out.shape
    pl = np.ascontiguousarray(pmap, dtype=np.int32)
    lib.orc_focuspix_apply(_p(out), w, h, black, _p(pl), len(pl), crop[0], crop[1], dual_iso)
    return out


def stripes_compute(img, black, white, frame_size, rng=None):
    lib = load_oracle()
    img = np.ascontiguousarray(img, dtype=np.uint16)
    h, w = img.shape
    if rng is None:
        rng = RandState()
        lib.orc_rand_seed(C.byref(rng), 1)
    coef = np.zeros(8, dtype=np.int32)
    needed = lib.orc_stripes_compute(_p(img), w, h, black, white, frame_size, C.byref(rng), _p(coef))
    return needed, coef


def stripes_apply(img, black, white, needed, coef):
    lib = load_oracle()
    out = np.ascontiguousarray(img, dtype=np.uint16).copy()
    h, w = out.shape
    coef = np.ascontiguousarray(coef, dtype=np.int32)
    lib.orc_stripes_apply(_p(out), out.size, w, black, white, int(needed), _p(coef))
    return out


def single_iso_chain(frames, black, white, frame_size, *, chroma_smooth_method=0, fix_bad_pixels=0,
                     fix_stripes=0, state=None):
    """process_frame's single-ISO stage order (main.c:966-997) over a list of unpacked frames.

    Per-clip state (bad-pixel list, stripe coefficients) comes from the first frame processed,
    exactly as in the reference.  Returns (list of frames, state).
    """
    state = state if state is not None else {}
    outs = []
    for fr in frames:
        img = np.ascontiguousarray(fr, dtype=np.uint16).copy()
        if fix_bad_pixels:
            if "badpix" not in state:
                state["badpix"] = badpix_detect(img, black, fix_bad_pixels == 2)
            img = badpix_apply(img, black, state["badpix"])
        if chroma_smooth_method:
            img = chroma_smooth(img, black, chroma_smooth_method)
        if fix_stripes:
            if "stripes" not in state:
                state["stripes"] = stripes_compute(img, black, white, frame_size)
            needed, coef = state["stripes"]
            img = stripes_apply(img, black, white, needed, coef)
        outs.append(img)
    return outs, state


# ---- LJ92 (oracle/orc_lj92.c) -------------------------------------------------------------------

def lj92_encode(tiled, depth=14):
    """Encode a (already quadrant-interleaved) uint16 image with the oracle's independent encoder."""
    lib = load_oracle()
    lib.orc_lj92_encode.restype = C.c_long
    tiled = np.ascontiguousarray(tiled, dtype=np.uint16)
    h, w = tiled.shape
    buf = np.zeros(w * h * 4 + 1024, dtype=np.uint8)
    n = lib.orc_lj92_encode(_p(tiled), w, h, depth, _p(buf), C.c_size_t(buf.size))
    if n < 0:
        raise RuntimeError("orc_lj92_encode failed")
    return buf[:n].copy()


def lj92_payload(img, depth=14):
    """VIDF payload of an LJ92 MLV frame: uint32 decoded size + stream of the interleaved image
    (inverse of main.c:656-668)."""
    h, w = img.shape
    tiled = np.concatenate([np.concatenate([img[0::2, 0::2], img[0::2, 1::2]], axis=1),
                            np.concatenate([img[1::2, 0::2], img[1::2, 1::2]], axis=1)], axis=0)
    stream = lj92_encode(tiled, depth)
    return np.concatenate([np.array([w * h * 2], dtype="<u4").view(np.uint8), stream])


def lj92_decode_payload(payload, w, h):
    """Oracle decode + de-interleave of a VIDF LJ92 payload -> uint16 [h, w]."""
    lib = load_oracle()
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    tiled = np.zeros(w * h, dtype=np.uint16)
    ww, hh, bb = C.c_int(), C.c_int(), C.c_int()
    rc = lib.orc_lj92_decode(C.c_void_p(payload.ctypes.data + 4), int(payload.size - 4), _p(tiled), w * h,
                             C.byref(ww), C.byref(hh), C.byref(bb))
    if rc != 0:
        raise RuntimeError(f"orc_lj92_decode failed: {rc}")
    out = np.zeros((h, w), dtype=np.uint16)
    lib.orc_lj92_untile(_p(tiled), _p(out), w, h)
    return out


# ---- pattern noise (oracle/orc_patternnoise.c) ---------------------------------------------------

def fix_pattern_noise(img, white):
    lib = load_oracle()
    out = np.ascontiguousarray(img, dtype=np.uint16).copy()
    h, w = out.shape
    lib.orc_fix_pattern_noise(_p(out), w, h, int(white))
    return out


# ---- dual ISO (oracle/orc_dualiso.c) -------------------------------------------------------------

class DisoState(C.Structure):
    _fields_ = [("raw2ev", C.c_void_p), ("ev2raw_0", C.c_void_p), ("lut_black", C.c_int), ("lut_white", C.c_int),
                ("fullres_curve", C.c_void_p), ("curve_black", C.c_int), ("amaze_fresh_tiles", C.c_int)]


class DisoInfo(C.Structure):
    _fields_ = [("rggb", C.c_int), ("is_bright", C.c_int * 4), ("white_dark", C.c_int), ("white_bright", C.c_int),
                ("white_darkened", C.c_int), ("a", C.c_double), ("b", C.c_double), ("corr_ev", C.c_double),
                ("overlap", C.c_double)]


def new_diso_state():
    st = DisoState()
    load_oracle().orc_diso_state_init(C.byref(st))
    return st


def cr2hdr20(img, black, white, *, interp_method=1, fullres=1, use_alias_map=1, chroma_smooth_method=0,
             fix_bad_pixels_mode=0, state=None, badpix_state=None, focus_map=None, crop=(0, 0)):
    """cr2hdr20_convert_data (hdr.c:1932-1957) on an unpacked frame.

    Returns (converted, out_frame, info); on converted == 1 the caller's black/white are x4.
    `badpix_state` is a dict holding the clip's bad-pixel list (detected on first use, cs.c:233-312).
    """
    lib = load_oracle()
    lib.orc_hdr_interpolate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]
    lib.orc_hdr_check.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    out = np.ascontiguousarray(img, dtype=np.uint16).copy()
    h, w = out.shape
    info = DisoInfo()
    if state is None:
        state = new_diso_state()
    if not lib.orc_hdr_check(_p(out), w, h, black, white):
        return 0, out, info
    if focus_map is not None and len(focus_map):
        out = focuspix_apply(out, black, focus_map, crop=crop, dual_iso=1)
    if fix_bad_pixels_mode:
        badpix_state = badpix_state if badpix_state is not None else {}
        if "list" not in badpix_state:
            badpix_state["list"] = badpix_detect(out, black, fix_bad_pixels_mode == 2, crop=crop)
        out = badpix_apply(out, black, badpix_state["list"], crop=crop, dual_iso=1)
    rc = lib.orc_hdr_interpolate(_p(out), w, h, black, interp_method, fullres, use_alias_map, chroma_smooth_method,
                                 C.byref(state), C.byref(info))
    if rc < 0:
        raise NotImplementedError("oracle: this dual-ISO interpolation method is not restated yet")
    return rc, out, info


# ---- dual-ISO preview + deflicker (oracle/orc_preview.c) -----------------------------------------

def hdr_preview(img, black, white, focus_map=None, crop=(0, 0)):
    """hdr_convert_data (hdr.c:40-227) -> (converted, frame); black/white x4 on success is the caller's job."""
    lib = load_oracle()
    lib.orc_hdr_preview.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t,
                                    C.c_int, C.c_int]
    out = np.ascontiguousarray(img, dtype=np.uint16).copy()
    h, w = out.shape
    fm = np.ascontiguousarray(focus_map, dtype=np.int32) if focus_map is not None and len(focus_map) else None
    rc = lib.orc_hdr_preview(_p(out), w, h, black, white, out.nbytes, _p(fm) if fm is not None else None,
                             len(fm) if fm is not None else 0, crop[0], crop[1])
    return rc, out


def deflicker(img, bpp, black, target):
    lib = load_oracle()
    lib.orc_deflicker.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p]
    img = np.ascontiguousarray(img, dtype=np.uint16)
    bias = (C.c_int * 2)()
    lib.orc_deflicker(_p(img), img.nbytes, bpp, black, target, bias)
    return bias[0], bias[1]


# ---- AMaZE demosaic (oracle/orc_amaze.c) ---------------------------------------------------------

def _amaze_planes(rawf):
    rawf = np.ascontiguousarray(rawf, dtype=np.float32)
    h, w = rawf.shape
    ws = w + 16                                     # hdr.c:969: rows are w+16 floats, zero-filled
    src = np.zeros((h, ws), np.float32)
    src[:, :w] = rawf
    outs = [np.full((h, ws), np.nan, np.float32) for _ in range(3)]
    return src, outs, h, w, ws


def amaze_demosaic(rawf, fresh_tiles=0):
    """orc_amaze_demosaic on an (h, w) float32 mosaic; returns red, green, blue (h, w) float32.
    Cells the algorithm never writes come back as NaN."""
    lib = load_oracle()
    lib.orc_amaze_demosaic.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4
    lib.orc_amaze_demosaic.restype = None
    src, outs, h, w, ws = _amaze_planes(rawf)
    lib.orc_amaze_demosaic(_p(src), _p(outs[0]), _p(outs[1]), _p(outs[2]), ws, w, h, int(fresh_tiles))
    return [o[:, :w].copy() for o in outs]


def ref_amaze_demosaic(rawf):
    """The compiled reference's amaze_demosaic_RT (amaze_demosaic_RT.c:113) on the same layout."""
    ref = load_ref()
    if ref is None:
        return None
    src, outs, h, w, ws = _amaze_planes(rawf)
    RowPtrs = C.POINTER(C.c_float) * h

    def rows(a):
        base = a.ctypes.data
        return RowPtrs(*[C.cast(base + r * ws * 4, C.POINTER(C.c_float)) for r in range(h)])

    ref.amaze_demosaic_RT.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4
    ref.amaze_demosaic_RT.restype = None
    ptrs = [rows(a) for a in [src] + outs]
    with quiet_stdout():
        ref.amaze_demosaic_RT(ptrs[0], ptrs[1], ptrs[2], ptrs[3], 0, 0, w, h)
    return [o[:, :w].copy() for o in outs]
