"""mlvfs_b200 -- B200 (sm_100a) implementation of MLVFS's per-frame raw path.

The product is ``libmlvfs_b200.so`` (CUDA kernels + C ABI, see include/mlvfs_b200.h) and the C host
code under ``mlvfs_b200/host``.  This Python module is only a thin ctypes binding of that ABI for
tests and bench.py; function names and argument meaning follow the reference's C interface
(dng.h, cs.h, stripes.h, hdr.h, patternnoise.h).

There is no CPU path: importing works anywhere, but ``lib()`` raises if the shared library is
missing, and every call fails if no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

from . import mlvformat
from .mlvformat import FrameHeaders

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmlvfs_b200.so")

OK, ERR_CUDA, ERR_ARG, ERR_UNSUPPORTED, ERR_NOMEM, ERR_NOT_DUAL_ISO = 0, -1, -2, -3, -4, -5


class Options(C.Structure):
    """``mlvb_options``: per-call snapshot of the processing fields of ``struct mlvfs`` (mlvfs.h:37-46)."""

    _fields_ = [
        ("chroma_smooth", C.c_int),
        ("fix_bad_pixels", C.c_int),
        ("fix_stripes", C.c_int),
        ("dual_iso", C.c_int),
        ("hdr_interpolation_method", C.c_int),
        ("hdr_no_fullres", C.c_int),
        ("hdr_no_alias_map", C.c_int),
        ("fix_pattern_noise", C.c_int),
        ("deflicker", C.c_int),
    ]


class FrameResult(C.Structure):
    _fields_ = [
        ("status", C.c_int),
        ("is_dual_iso", C.c_int),
        ("black_level", C.c_int32),
        ("white_level", C.c_int32),
        ("exposure_bias", C.c_int32 * 2),
    ]


class StripesCorrection(C.Structure):
    """``struct stripes_correction`` (stripes.h:30-36)."""


StripesCorrection._fields_ = [
    ("next", C.POINTER(StripesCorrection)),
    ("mlv_filename", C.c_char_p),
    ("correction_needed", C.c_int),
    ("coeffficients", C.c_int * 8),
]

# every symbol include/mlvfs_b200.h declares (tests/test_abi_symbols.py checks the header against this)
ABI_SYMBOLS = [
    "mlvb_context_create", "mlvb_context_destroy", "mlvb_default_context", "mlvb_device_count",
    "mlvb_host_alloc", "mlvb_host_free", "mlvb_host_pool_trim", "mlvb_process_frame", "mlvb_submit", "mlvb_wait",
    "mlvb_process_frames", "mlvb_process_batch_device", "mlvb_reset_clip_state", "mlvb_seed_dither", "mlvb_get_stripes",
    "mlvb_get_bad_pixels", "mlvb_launch_count", "mlvb_path_count", "mlvb_profile_begin", "mlvb_profile_end",
    "dng_get_image_data", "dng_get_image_size", "get_image_data", "get_raw2evf", "get_raw2ev", "get_ev2raw",
    "chroma_smooth", "fix_bad_pixels", "fix_focus_pixels", "free_focus_pixel_maps",
    "stripes_get_correction", "stripes_new_correction", "stripes_free_corrections",
    "stripes_compute_correction", "stripes_apply_correction",
    "fix_pattern_noise", "hdr_convert_data", "cr2hdr20_convert_data", "amaze_demosaic_RT",
]

_lib = None


def build(verbose=False):
    """Compile libmlvfs_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j8"] + ([] if verbose else ["-s"])
    subprocess.check_call(cmd)


def lib():
    """The loaded C ABI.  Raises if the CUDA library is not built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                               "mlvfs_b200 has no CPU path.")
        L = C.CDLL(LIB_PATH)
        vp, sz = C.c_void_p, C.c_size_t
        L.mlvb_context_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
        L.mlvb_context_destroy.argtypes = [vp]
        L.mlvb_default_context.restype = vp
        L.mlvb_host_alloc.restype = vp
        L.mlvb_host_alloc.argtypes = [sz]
        L.mlvb_host_free.argtypes = [vp]
        L.mlvb_process_frame.argtypes = [vp, vp, vp, sz, vp, C.c_char_p, vp, vp]
        L.mlvb_submit.restype = C.c_int64
        L.mlvb_submit.argtypes = [vp, vp, vp, sz, vp, C.c_char_p, vp]
        L.mlvb_wait.argtypes = [vp, C.c_int64, vp]
        L.mlvb_process_frames.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_char_p, vp, vp]
        L.mlvb_host_pool_trim.argtypes = []
        L.mlvb_process_batch_device.argtypes = [vp, vp, vp, C.c_char_p, vp, sz, sz, vp, sz, C.c_int, vp]
        L.mlvb_reset_clip_state.argtypes = [vp]
        L.mlvb_seed_dither.argtypes = [vp, C.c_uint]
        L.mlvb_get_stripes.argtypes = [vp, C.c_char_p, vp, vp]
        L.mlvb_get_bad_pixels.argtypes = [vp, C.c_uint64, C.c_int, vp, C.c_int]
        L.mlvb_launch_count.restype = C.c_uint64
        L.mlvb_launch_count.argtypes = [vp]
        L.mlvb_path_count.restype = C.c_uint64
        L.mlvb_path_count.argtypes = [vp, C.c_int]
        L.mlvb_profile_begin.argtypes = [vp]
        L.mlvb_profile_end.argtypes = [vp, vp, vp, C.c_int]
        L.dng_get_image_data.restype = sz
        L.dng_get_image_data.argtypes = [vp, vp, vp, C.c_long, sz]
        L.dng_get_image_size.restype = sz
        L.dng_get_image_size.argtypes = [vp]
        L.get_raw2ev.restype = C.POINTER(C.c_int)
        L.get_ev2raw.restype = C.POINTER(C.c_int)
        L.get_raw2evf.restype = C.POINTER(C.c_double)
        L.chroma_smooth.argtypes = [vp, vp, C.c_int]
        L.fix_bad_pixels.argtypes = [vp, vp, C.c_int, C.c_int]
        L.fix_focus_pixels.argtypes = [vp, vp, C.c_int]
        L.stripes_get_correction.restype = C.POINTER(StripesCorrection)
        L.stripes_get_correction.argtypes = [C.c_char_p]
        L.stripes_new_correction.restype = C.POINTER(StripesCorrection)
        L.stripes_new_correction.argtypes = [C.c_char_p]
        L.stripes_compute_correction.argtypes = [vp, vp, vp, C.c_long, sz]
        L.stripes_apply_correction.argtypes = [vp, vp, vp, C.c_long, sz]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One ``mlvb_context`` (one GPU).  Mirrors the frame-request side of process_frame (main.c:908)."""

    def __init__(self, device=0, slots=4):
        self._h = C.c_void_p()
        rc = lib().mlvb_context_create(device, slots, C.byref(self._h))
        if rc != OK:
            raise RuntimeError(f"mlvb_context_create(device={device}) failed with {rc} (no CUDA device? no CPU path)")
        self.device = device

    def close(self):
        if self._h:
            lib().mlvb_context_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def process_frame(self, hdr, payload, opts, mlv_filename, out=None):
        """payload: numpy uint8/uint16 array holding the VIDF payload.  Returns (uint16 [h,w], FrameResult)."""
        w, h = hdr.rawi_hdr.xRes, hdr.rawi_hdr.yRes
        payload = np.ascontiguousarray(payload)
        if out is None:
            out = np.empty((h, w), dtype=np.uint16)
        res = FrameResult()
        rc = lib().mlvb_process_frame(self._h, C.byref(hdr), _ptr(payload), payload.nbytes, C.byref(opts),
                                      mlv_filename.encode(), _ptr(out), C.byref(res))
        if rc != OK:
            raise RuntimeError(f"mlvb_process_frame failed with {rc}")
        return out, res

    def process_frames(self, hdrs, payload_ptrs, payload_bytes, opts, mlv_filename, dst_ptrs):
        """mlvb_process_frames: a host batch of frames of one clip.  hdrs: list of FrameHeaders; payload_ptrs /
        dst_ptrs: host addresses (ints); returns (rc, [FrameResult])."""
        n = len(hdrs)
        H = (FrameHeaders * n)(*hdrs)
        P = (C.c_void_p * n)(*payload_ptrs)
        B = (C.c_size_t * n)(*payload_bytes)
        D = (C.c_void_p * n)(*dst_ptrs)
        R = (FrameResult * n)()
        rc = lib().mlvb_process_frames(self._h, n, H, P, B, C.byref(opts), mlv_filename.encode(), D, R)
        return rc, list(R)

    def submit(self, hdr, payload_ptr, payload_bytes, opts, mlv_filename, dst_ptr):
        return lib().mlvb_submit(self._h, C.byref(hdr), payload_ptr, payload_bytes, C.byref(opts),
                                 mlv_filename.encode(), dst_ptr)

    def wait(self, ticket):
        res = FrameResult()
        rc = lib().mlvb_wait(self._h, ticket, C.byref(res))
        return rc, res

    def process_batch_device(self, hdr, opts, mlv_filename, d_payload, payload_stride, payload_bytes, d_out,
                             out_stride_px, nframes, stream=0):
        rc = lib().mlvb_process_batch_device(self._h, C.byref(hdr), C.byref(opts), mlv_filename.encode(),
                                             C.c_void_p(d_payload), payload_stride, payload_bytes,
                                             C.c_void_p(d_out), out_stride_px, nframes, C.c_void_p(stream))
        if rc != OK:
            raise RuntimeError(f"mlvb_process_batch_device failed with {rc}")

    def reset_clip_state(self):
        lib().mlvb_reset_clip_state(self._h)

    def seed_dither(self, seed=1):
        lib().mlvb_seed_dither(self._h, seed)

    def get_stripes(self, mlv_filename):
        needed = C.c_int()
        coef = (C.c_int * 8)()
        rc = lib().mlvb_get_stripes(self._h, mlv_filename.encode(), C.byref(needed), coef)
        if rc < 0:
            return None
        return needed.value, np.array(list(coef), dtype=np.int32)

    def get_bad_pixels(self, file_guid, aggressive):
        n = lib().mlvb_get_bad_pixels(self._h, file_guid, int(aggressive), None, 0)
        if n < 0:
            return None
        xy = np.zeros((max(n, 1), 2), dtype=np.int32)
        lib().mlvb_get_bad_pixels(self._h, file_guid, int(aggressive), _ptr(xy), n)
        return xy[:n]

    def launch_count(self):
        return int(lib().mlvb_launch_count(self._h))

    def path_count(self, which):
        return int(lib().mlvb_path_count(self._h, int(which)))

    STAGES = ["unpack", "pixfix", "chroma", "stripes", "pattern", "dualiso", "lj92", "other"]

    def profile_begin(self):
        lib().mlvb_profile_begin(self._h)

    def profile_end(self):
        """-> {stage: (total_ms, spans)} for the stages that ran since profile_begin."""
        ms = (C.c_float * 8)()
        cnt = (C.c_int * 8)()
        lib().mlvb_profile_end(self._h, ms, cnt, 8)
        return {self.STAGES[i]: (float(ms[i]), int(cnt[i])) for i in range(8) if cnt[i]}

    @classmethod
    def default(cls):
        """Wrapper around the process-wide context the drop-in symbols use (not owned)."""
        self = cls.__new__(cls)
        self._h = C.c_void_p(lib().mlvb_default_context())
        if not self._h:
            raise RuntimeError("mlvb_default_context() failed (no CUDA device? no CPU path)")
        self.device = -1
        self.close = lambda: None
        return self


class PinnedBuffer:
    """Pinned host memory from mlvb_host_alloc, viewed as a numpy array."""

    def __init__(self, nbytes, dtype=np.uint8):
        self.ptr = lib().mlvb_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError("mlvb_host_alloc failed")
        self.nbytes = nbytes
        buf = (C.c_uint8 * nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype)

    def free(self):
        if self.ptr:
            self.array = None
            lib().mlvb_host_free(self.ptr)
            self.ptr = None


# ---- reference-named drop-in calls on host numpy buffers (default context) ----------------------

def dng_get_image_data(hdr, packed_words, offset=0, max_size=None):
    """dng.h:31 -- unpack; packed_words must start at the word holding the first requested pixel."""
    w, h = hdr.rawi_hdr.xRes, hdr.rawi_hdr.yRes
    max_size = w * h * 2 if max_size is None else max_size
    packed_words = np.ascontiguousarray(packed_words, dtype=np.uint16)
    out = np.zeros(max_size, dtype=np.uint8)
    n = lib().dng_get_image_data(C.byref(hdr), _ptr(packed_words), _ptr(out), offset, max_size)
    return n, out


def chroma_smooth(hdr, image, method):
    """cs.h:27 -- in place on a uint16 [h,w] array."""
    assert image.dtype == np.uint16 and image.flags.c_contiguous
    lib().chroma_smooth(C.byref(hdr), _ptr(image), method)
    return image


def fix_bad_pixels(hdr, image, aggressive, dual_iso):
    assert image.dtype == np.uint16 and image.flags.c_contiguous
    lib().fix_bad_pixels(C.byref(hdr), _ptr(image), int(aggressive), int(dual_iso))
    return image


def fix_focus_pixels(hdr, image, dual_iso):
    assert image.dtype == np.uint16 and image.flags.c_contiguous
    lib().fix_focus_pixels(C.byref(hdr), _ptr(image), int(dual_iso))
    return image


def stripes_compute_correction(hdr, image, mlv_filename):
    L = lib()
    corr = L.stripes_get_correction(mlv_filename.encode())
    if not corr:
        corr = L.stripes_new_correction(mlv_filename.encode())
        L.stripes_compute_correction(C.byref(hdr), corr, _ptr(image), 0, image.size)
    return corr


def stripes_apply_correction(hdr, corr, image, offset=0):
    lib().stripes_apply_correction(C.byref(hdr), corr, _ptr(image), offset, image.size)
    return image


def amaze_demosaic(rawf):
    """The drop-in ``amaze_demosaic_RT`` (reference amaze_demosaic_RT.c:113) on an (h, w) float32 RGGB mosaic,
    called exactly like hdr.c:1040 calls it: arrays of row pointers, rows of w+16 floats.  Returns red, green,
    blue as (h, w) float32."""
    rawf = np.ascontiguousarray(rawf, dtype=np.float32)
    h, w = rawf.shape
    ws = w + 16
    src = np.zeros((h, ws), np.float32)
    src[:, :w] = rawf
    outs = [np.full((h, ws), np.nan, np.float32) for _ in range(3)]
    Rows = C.POINTER(C.c_float) * h

    def rows(a):
        return Rows(*[C.cast(a.ctypes.data + r * ws * 4, C.POINTER(C.c_float)) for r in range(h)])

    L = lib()
    L.amaze_demosaic_RT.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4
    L.amaze_demosaic_RT.restype = None
    ptrs = [rows(a) for a in [src] + outs]
    L.amaze_demosaic_RT(ptrs[0], ptrs[1], ptrs[2], ptrs[3], 0, 0, w, h)
    return [o[:, :w].copy() for o in outs]
