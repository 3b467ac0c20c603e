"""GPU parity: full dual-ISO conversion (--dual-iso --mean23), BASELINE config 3 options.

Statistics (row fields, white levels, exposure match) are integer-exact; the per-pixel blends are fp64
with CUDA's log2/cos, so the frame is a tolerance stage: <= 1 DN on the 16-bit output (PSNR reported)."""
import ctypes as C

import numpy as np
import pytest

import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F, synth

pytestmark = pytest.mark.gpu

TOL_DN = 1     # north_star tolerance for the floating-point dual-ISO stages


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10 * np.log10(65535.0 ** 2 / mse)


def _check(got, want, label):
    from conftest import parity_record
    rec = parity_record(label, got, want, TOL_DN)
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert rec["max_abs_diff_dn"] <= TOL_DN, f"{label}: max diff {d.max()} DN at {np.unravel_index(d.argmax(), d.shape)}"


@pytest.mark.parametrize("w,h,cs,alias,badpix,fullres", [
    (640, 360, 0, 0, 0, 1), (640, 360, 5, 1, 0, 1), (640, 362, 3, 1, 2, 1), (1920, 1080, 5, 1, 0, 1), (384, 216, 3, 1, 0, 0)])
def test_cr2hdr20_dropin_matches_oracle(fresh_ctx, oracle, w, h, cs, alias, badpix, fullres):
    img = synth.make_frame(w, h, 0, dual_iso=True, hot_cold=True, bad_density=1e-4)
    hdr = F.make_frame_headers(w, h, file_guid=0xA100 + cs * 64 + alias * 16 + badpix * 4 + fullres)
    rc, want, info = oracle.cr2hdr20(img, 2048, 15000, interp_method=1, fullres=fullres, use_alias_map=alias,
                                     chroma_smooth_method=cs, fix_bad_pixels_mode=badpix)
    assert rc == 1
    got = img.copy()
    L = M.lib()
    L.cr2hdr20_convert_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    r = L.cr2hdr20_convert_data(C.byref(hdr), got.ctypes.data_as(C.c_void_p), 1, fullres, alias, cs, badpix)
    assert r == 1
    assert hdr.rawi_hdr.raw_info.black_level == 8192 and hdr.rawi_hdr.raw_info.white_level == 60000
    _check(got, want, f"cr2hdr20 {w}x{h} cs{cs} alias{alias} badpix{badpix} fullres{fullres}")


def test_dual_iso_row_phases_and_gbrg(fresh_ctx, oracle):
    w, h = 480, 272
    base = synth.make_frame(w, h + 3, 3, dual_iso=True)
    L = M.lib()
    L.cr2hdr20_convert_data.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    for k, img in enumerate([base[2:h + 2], base[1:h + 1], base[3:h + 3]]):
        img = np.ascontiguousarray(img)
        hdr = F.make_frame_headers(w, h, file_guid=0xA200 + k)
        rc, want, info = oracle.cr2hdr20(img, 2048, 15000, interp_method=1, chroma_smooth_method=3)
        got = img.copy()
        r = L.cr2hdr20_convert_data(C.byref(hdr), got.ctypes.data_as(C.c_void_p), 1, 1, 1, 3, 0)
        assert r == rc, k
        _check(got, want, f"phase/cfa case {k} (rggb={info.rggb}, pattern={list(info.is_bright)})")


def test_plain_footage_is_not_converted(fresh_ctx, oracle):
    """--dual-iso on ordinary footage: detection fails, frame keeps the horizontal bad-pixel repairs, then
    process_frame repairs again in 2-D and skips chroma smoothing (main.c:956-978)."""
    w, h = 640, 360
    hdr = F.make_frame_headers(w, h, file_guid=0xA300)
    ri = hdr.rawi_hdr.raw_info
    img = synth.make_frame(w, h, 1, hot_cold=True, bad_density=1e-4)
    rc, mid, _ = oracle.cr2hdr20(img, 2048, 15000, interp_method=1, fix_bad_pixels_mode=1)
    assert rc == 0
    lst = oracle.badpix_detect(img, 2048, 0)           # the map was detected on the untouched frame
    want = oracle.badpix_apply(mid, 2048, lst)
    o = M.Options(dual_iso=2, hdr_interpolation_method=1, fix_bad_pixels=1, chroma_smooth=3)
    out, res = fresh_ctx.process_frame(hdr, synth.pack_bits(img), o, "plain.MLV")
    assert res.is_dual_iso == 0 and res.black_level == 2048
    assert np.array_equal(out, want)


def test_dual_iso_through_process_frame(fresh_ctx, oracle):
    """BASELINE config 3 options through the fused entry: --dual-iso --mean23 --cs5x5 (alias map on)."""
    w, h = 960, 384
    hdr = F.make_frame_headers(w, h, file_guid=0xA400)
    o = M.Options(dual_iso=2, hdr_interpolation_method=1, chroma_smooth=5)
    st = oracle.new_diso_state()
    for i in range(2):
        img = synth.make_frame(w, h, i, dual_iso=True)
        rc, want, _ = oracle.cr2hdr20(img, 2048, 15000, interp_method=1, chroma_smooth_method=5, state=st)
        out, res = fresh_ctx.process_frame(hdr, synth.pack_bits(img), o, "c3.MLV")
        assert rc == 1 and res.is_dual_iso == 1 and res.black_level == 8192 and res.white_level == 60000
        _check(out, want, f"process_frame dual ISO frame {i}")


@pytest.mark.parametrize("method,cs", [(1, 5), (0, 0)])
def test_dual_iso_device_batch_matches_per_frame(fresh_ctx, method, cs):
    """The device-resident batch entry runs the frames after the first on a few worker lanes (own stream, own
    scratch, own host thread); every frame must equal what the per-frame call gives (bit for bit: same kernels,
    same statistics), in any interleaving."""
    torch = pytest.importorskip("torch")
    w, h, n = 640, 384, 7
    hdr = F.make_frame_headers(w, h)
    frames = [synth.make_frame(w, h, i, dual_iso=True, hot_cold=True) for i in range(n)]
    o = M.Options(dual_iso=2, hdr_interpolation_method=method, chroma_smooth=cs, fix_bad_pixels=1)
    want = []
    for i, img in enumerate(frames):
        out, res = fresh_ctx.process_frame(hdr, synth.pack_bits(img), o, "dbatch.MLV")
        assert res.status == 0 and res.is_dual_iso == 1
        want.append(out.copy())
    packed = np.stack([synth.pack_bits(f) for f in frames])
    stride = packed.shape[1] * 2
    d_in = torch.from_numpy(packed.view(np.int16)).cuda()
    d_out = torch.zeros((n, h * w), dtype=torch.int16, device="cuda")
    for rep in range(2):
        d_out.zero_()
        fresh_ctx.process_batch_device(hdr, o, "dbatch.MLV", d_in.data_ptr(), stride, stride, d_out.data_ptr(), h * w, n,
                                       torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(np.uint16).reshape(n, h, w)
        for i in range(n):
            assert np.array_equal(got[i], want[i]), (rep, i, int(np.count_nonzero(got[i] != want[i])))


def test_dual_iso_pixel_fix_long_rows(fresh_ctx, oracle):
    """Rows with many bad pixels are staged in shared memory and repaired as independent runs (entries more than
    3 columns apart) on different lanes; runs of neighbouring entries stay sequential.  Bit-exact against the
    list-order walk of the reference (cs.c:314-330 with dual_iso = 1), including entries at the row ends."""
    w, h = 2048, 96
    hdr = F.make_frame_headers(w, h, file_guid=0xB0B0)
    img = synth.make_frame(w, h, 3, dual_iso=True)
    rng = np.random.default_rng(5)
    # ~12 % hot pixels with runs at distance 1, 2, 3 and 4 (the last one is the first independent distance)
    hot = rng.random((h, w)) < 0.06
    for d in (1, 2, 3, 4):
        seed = rng.random((h, w)) < 0.01
        hot |= seed | np.roll(seed, d, axis=1)
    hot[:, :8] |= rng.random((h, 8)) < 0.3           # edge rules at both row ends
    hot[:, -8:] |= rng.random((h, 8)) < 0.3
    img[hot] = 16000
    want_list = oracle.badpix_detect(img, 2048, 1)
    per_row = np.bincount(want_list[:, 1], minlength=h)
    assert per_row.max() >= 128, "no row reaches the long-row path"
    first = M.fix_bad_pixels(hdr, img.copy(), 1, 0)              # detects the map (2-D interpolator, level schedule)
    assert np.array_equal(first, oracle.badpix_apply(img, 2048, want_list))
    img2 = synth.make_frame(w, h, 4, dual_iso=True)
    img2[hot] = 15900
    got = M.fix_bad_pixels(hdr, img2.copy(), 1, 1)               # same map, horizontal interpolator
    want = oracle.badpix_apply(img2, 2048, want_list, dual_iso=1)
    assert np.array_equal(got, want), int(np.count_nonzero(got != want))


@pytest.mark.parametrize("content", ["frame", "steps", "noise", "flat", "ramp"])
def test_dual_iso_dense_rows_are_walked_by_the_whole_warp(fresh_ctx, oracle, content):
    """--really-bad-pix on a dual-ISO frame flags (nearly) every pixel of the bright rows: each such row is one
    recurrence thousands of steps long (cs.c:87-109 applied in list order).  The CUDA path walks it in 32 speculative
    pieces and re-walks those whose assumed start state was wrong, so the result must equal the serial walk whatever
    the content: a normal frame (the speculation converges), two-column steps (the repaired value is carried along
    the row: the speculation fails and pieces are redone), full-range noise, flat rows (sum == 0: copies) and a ramp.
    Three windows of 1024 pixels per row, the last one partial."""
    w, h = 2600, 48
    hdr = F.make_frame_headers(w, h, file_guid=0xD15E + hash(content) % 1000)
    img = synth.make_frame(w, h, 7, dual_iso=True)
    want_list = oracle.badpix_detect(img, 2048, 1)
    per_row = np.bincount(want_list[:, 1], minlength=h)
    assert per_row.max() >= 2000, f"no dense row: the longest has {per_row.max()} entries"
    first = M.fix_bad_pixels(hdr, img.copy(), 1, 1)                  # detects the map and repairs the frame it came from
    assert np.array_equal(first, oracle.badpix_apply(img, 2048, want_list, dual_iso=1))
    rng = np.random.default_rng(11)
    x = np.arange(w)[None, :]
    if content == "frame":
        img2 = synth.make_frame(w, h, 8, dual_iso=True)
    elif content == "steps":
        img2 = (2200 + 9000 * ((x // 2) % 2) + rng.integers(0, 3, (h, w))).astype(np.uint16)
    elif content == "noise":
        img2 = rng.integers(1900, 16383, (h, w)).astype(np.uint16)
    elif content == "flat":
        img2 = np.full((h, w), 5000, np.uint16)
        img2[:, ::97] = 9000
    else:
        img2 = (2100 + (x * 5) % 14000 + rng.integers(0, 2, (h, w))).astype(np.uint16)
    got = M.fix_bad_pixels(hdr, img2.copy(), 1, 1)
    want = oracle.badpix_apply(img2, 2048, want_list, dual_iso=1)
    assert np.array_equal(got, want), (content, int(np.count_nonzero(got != want)))


def test_dual_iso_frames_in_flight_match_one_at_a_time(fresh_ctx):
    """mlvb_submit hands dual-ISO frames to its submit workers once the clip's state exists; several frames in flight
    (pinned and pageable buffers mixed) must give exactly the frames that one-at-a-time processing gives."""
    w, h, n = 640, 384, 7
    hdr = F.make_frame_headers(w, h, file_guid=0xA5A5)
    o = M.Options(dual_iso=2, hdr_interpolation_method=1, chroma_smooth=3, fix_bad_pixels=1)
    frames = [synth.make_frame(w, h, i, dual_iso=True, hot_cold=True) for i in range(n)]
    packed = [synth.pack_bits(f) for f in frames]
    want = []
    for p in packed:
        out, res = fresh_ctx.process_frame(hdr, p, o, "inflight.MLV")
        assert res.status == 0 and res.is_dual_iso == 1
        want.append(out.copy())
    nbytes = packed[0].size * 2
    pin_in = M.PinnedBuffer(n * nbytes)
    pin_in.array[:] = np.concatenate([p.view(np.uint8).reshape(-1) for p in packed])
    pin_out = [M.PinnedBuffer(w * h * 2) for _ in range(n)]
    pageable_out = np.zeros((h, w), np.uint16)
    try:
        depth = 3                                                # the fixture's context has 4 slots: never ask for a 5th
        for rep in range(2):
            pending = []

            def finish(i, t):
                rc, res = fresh_ctx.wait(t)
                assert rc == 0 and res.is_dual_iso == 1 and res.black_level == 8192
                got = pageable_out if i == 3 else pin_out[i].array.view(np.uint16).reshape(h, w)
                assert np.array_equal(got, want[i]), (rep, i)

            for i in range(n):
                if len(pending) == depth:
                    finish(*pending.pop(-1 if i % 2 else 0))     # not always the oldest
                dst = pageable_out.ctypes.data if i == 3 else pin_out[i].ptr
                src = packed[i].ctypes.data if i == 5 else pin_in.ptr + i * nbytes
                t = fresh_ctx.submit(hdr, C.c_void_p(src), nbytes, o, "inflight.MLV", C.c_void_p(dst))
                assert t >= 0
                pending.append((i, t))
            while pending:
                finish(*pending.pop())
    finally:
        pin_in.free()
        for p in pin_out:
            p.free()
