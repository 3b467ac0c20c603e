"""GPU parity: --dual-iso-preview (hdr.c:40-227) and --deflicker (main.c:895-906) vs the oracle."""
import ctypes as C

import numpy as np
import pytest

import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F, synth

pytestmark = pytest.mark.gpu

# The least-squares a, b come from exact integer histograms through the same fp64 operation order on the host
# (hdr.c:154-183) and the per-pixel scaling is fp64 without FMA contraction: the preview is bit-exact.
PREVIEW_TOL_DN = 0


@pytest.mark.parametrize("w,h,shift,white", [(640, 360, 0, 15000), (642, 362, 1, 15000), (640, 360, 2, 12000),
                                             (1920, 1080, 3, 15000)])
def test_hdr_convert_data_dropin(fresh_ctx, oracle, w, h, shift, white):
    base = synth.make_frame(w, h + 4, 0, dual_iso=True, white=white)
    img = np.ascontiguousarray(base[shift:shift + h])
    if white == 12000:
        img[100:140, 200:300] = 0
    hdr = F.make_frame_headers(w, h, white=white)
    rc, want = oracle.hdr_preview(img, 2048, white)
    assert rc == 1
    got = img.copy()
    L = M.lib()
    L.hdr_convert_data.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_size_t]
    assert L.hdr_convert_data(C.byref(hdr), got.ctypes.data_as(C.c_void_p), 0, got.nbytes) == 1
    assert hdr.rawi_hdr.raw_info.black_level == 8192 and hdr.rawi_hdr.raw_info.white_level == white * 4
    from conftest import parity_record
    rec = parity_record(f"hdr_convert_data preview {w}x{h} shift{shift} white{white}", got, want, PREVIEW_TOL_DN)
    assert rec["max_abs_diff_dn"] <= PREVIEW_TOL_DN
    plain = synth.make_frame(w, h, 1)
    keep = plain.copy()
    assert L.hdr_convert_data(C.byref(F.make_frame_headers(w, h, white=white)), plain.ctypes.data_as(C.c_void_p), 0, plain.nbytes) == 0
    assert np.array_equal(plain, keep)


def test_preview_and_deflicker_through_process_frame(fresh_ctx, oracle):
    w, h = 640, 360
    hdr = F.make_frame_headers(w, h)
    img = synth.make_frame(w, h, 2, dual_iso=True)
    rc, want = oracle.hdr_preview(img, 2048, 15000)
    bias = oracle.deflicker(img, 14, 2048, 3000)
    out, res = fresh_ctx.process_frame(hdr, synth.pack_bits(img), M.Options(dual_iso=1, deflicker=3000), "pv.MLV")
    assert res.is_dual_iso == 1 and res.black_level == 8192 and res.white_level == 60000
    assert (res.exposure_bias[0], res.exposure_bias[1]) == bias
    assert np.abs(out.astype(np.int32) - want.astype(np.int32)).max() <= PREVIEW_TOL_DN
    # deflicker alone leaves the pixels untouched
    plain = synth.make_frame(w, h, 4)
    out, res = fresh_ctx.process_frame(hdr, synth.pack_bits(plain), M.Options(deflicker=4500), "pv2.MLV")
    assert np.array_equal(out, plain)
    assert (res.exposure_bias[0], res.exposure_bias[1]) == oracle.deflicker(plain, 14, 2048, 4500)
