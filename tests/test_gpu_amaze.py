"""GPU parity: AMaZE demosaic (amaze_demosaic_RT drop-in) and the dual-ISO --amaze-edge path, BASELINE config 4 options.

The AMaZE planes are compared bit for bit (uint32 view of the floats) with the oracle's independent-tile mode, which
tests/test_oracle_vs_ref.py shows equal to the compiled reference for widths that are multiples of 128.  The whole
dual-ISO frame is a tolerance stage like the mean23 path (fp64 log2/cos in the blends): <= 1 DN, PSNR reported."""
import ctypes as C

import numpy as np
import pytest

import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F, synth

pytestmark = pytest.mark.gpu
TOL_DN = 1


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10 * np.log10(65535.0 ** 2 / mse)


@pytest.mark.parametrize("w,h", [(160, 160), (256, 206), (384, 270), (640, 360), (1920, 544)])
def test_amaze_planes_bit_exact(fresh_ctx, oracle, w, h):
    raw = synth.amaze_test_mosaic(w, h, w * 7 + h)
    want = oracle.amaze_demosaic(raw, fresh_tiles=1)
    got = M.amaze_demosaic(raw)
    for name, a, b in zip("RGB", got, want):
        ne = a.view(np.uint32) != b.view(np.uint32)
        assert not ne.any(), f"{name}: {int(ne.sum())} floats differ, first at {np.argwhere(ne)[0]}, max |d| {np.nanmax(np.abs(a - b))}"
    overrun = any(144 < h - top < 160 for top in range(-16, h, 128))   # the reference's bottom-border overrun
    if w % 128 == 0 and not overrun:
        ref_like = oracle.amaze_demosaic(raw, fresh_tiles=0)        # the reference's sequential tile walk
        for a, b in zip(got, ref_like):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def _check(got, want, label):
    from conftest import parity_record
    parity_record(label, got, want, TOL_DN)
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= TOL_DN, f"{label}: max diff {d.max()} DN at {np.unravel_index(d.argmax(), d.shape)}"


@pytest.mark.parametrize("w,h,cs,alias,badpix,fullres", [
    (640, 360, 0, 0, 0, 1), (640, 362, 3, 1, 2, 1), (384, 216, 5, 1, 0, 0), (1920, 1080, 0, 1, 2, 1)])
def test_cr2hdr20_amaze_dropin_matches_oracle(fresh_ctx, oracle, w, h, cs, alias, badpix, fullres):
    img = synth.make_frame(w, h, 0, dual_iso=True, hot_cold=True, bad_density=1e-4)
    hdr = F.make_frame_headers(w, h, file_guid=0xC100 + cs * 64 + alias * 16 + badpix * 4 + fullres)
    rc, want, info = oracle.cr2hdr20(img, 2048, 15000, interp_method=0, fullres=fullres, use_alias_map=alias,
                                     chroma_smooth_method=cs, fix_bad_pixels_mode=badpix)
    assert rc == 1
    got = img.copy()
    L = M.lib()
    L.cr2hdr20_convert_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    r = L.cr2hdr20_convert_data(C.byref(hdr), got.ctypes.data_as(C.c_void_p), 0, fullres, alias, cs, badpix)
    assert r == 1
    assert hdr.rawi_hdr.raw_info.black_level == 8192 and hdr.rawi_hdr.raw_info.white_level == 60000
    _check(got, want, f"cr2hdr20 amaze {w}x{h} cs{cs} alias{alias} badpix{badpix} fullres{fullres}")


def test_amaze_through_process_frame(fresh_ctx, oracle):
    """BASELINE config 4 options through the fused entry: --dual-iso --amaze-edge --alias-map --really-bad-pix."""
    w, h = 768, 392
    hdr = F.make_frame_headers(w, h, file_guid=0xC400)
    o = M.Options(dual_iso=2, hdr_interpolation_method=0, fix_bad_pixels=2)
    st = oracle.new_diso_state()
    bp = {}
    for i in range(2):
        img = synth.make_frame(w, h, i, dual_iso=True, hot_cold=True, bad_density=1e-4)
        rc, want, _ = oracle.cr2hdr20(img, 2048, 15000, interp_method=0, fix_bad_pixels_mode=2, state=st, badpix_state=bp)
        out, res = fresh_ctx.process_frame(hdr, synth.pack_bits(img), o, "c4.MLV")
        assert rc == 1 and res.is_dual_iso == 1 and res.black_level == 8192 and res.white_level == 60000
        _check(out, want, f"process_frame amaze frame {i}")
