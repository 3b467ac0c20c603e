"""GPU parity at the BASELINE.json geometries that carry size-dependent code paths: C3 3840x1536 (--dual-iso --mean23
--cs5x5), C4 5760x3240 (--dual-iso --amaze-edge --alias-map --really-bad-pix: AMaZE tile grid with partial bottom
tiles, long-row pixel repair in sliding windows), C5 3840x2160 LJ92 (subsequence / block counts).  The oracle here is
the UNMODIFIED reference's process_frame (oracle/_ref) on the same synthetic MLV, run in a fresh process; our side
goes through mlvb_process_frame.  Integer stages bit-exact, the dual-ISO frames within 1 DN (north star), max |diff|
and PSNR recorded in gpurun_out/parity_r02.json."""
import os

import numpy as np
import pytest

import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F, synth

from conftest import parity_record, reference_frames

pytestmark = pytest.mark.gpu


def run_ours(ctx, hdr, payloads, opts, clip):
    outs, results = [], []
    for p in payloads:
        out, res = ctx.process_frame(hdr, p, M.Options(**opts), clip)
        assert res.status == 0
        outs.append(out.copy())
        results.append(res)
    return outs, results


def test_c3_geometry_dual_iso_mean23_cs5x5(fresh_ctx, tmp_path):
    w, h, n = 3840, 1536, 2
    opts = dict(dual_iso=2, hdr_interpolation_method=1, chroma_smooth=5)
    hdr, frames = synth.make_clip(str(tmp_path / "C3.MLV"), w, h, n, variant=dict(dual_iso=True))
    want = reference_frames(tmp_path, "C3.MLV", n, opts)
    got, res = run_ours(fresh_ctx, hdr, [synth.pack_bits(f) for f in frames], opts, "C3.MLV")
    for i in range(n):
        assert res[i].is_dual_iso == 1 and res[i].black_level == 8192 and res[i].white_level == 60000
        rec = parity_record(f"C3 3840x1536 dual-iso mean23 cs5x5 frame {i}", got[i], want[i], 1)
        assert rec["max_abs_diff_dn"] <= 1


def test_c4_geometry_dual_iso_amaze_alias_really_bad_pix(fresh_ctx, tmp_path):
    w, h, n = 5760, 3240, 1
    opts = dict(dual_iso=2, hdr_interpolation_method=0, fix_bad_pixels=2)
    hdr, frames = synth.make_clip(str(tmp_path / "C4.MLV"), w, h, n, variant=dict(dual_iso=True, hot_cold=True))
    want = reference_frames(tmp_path, "C4.MLV", n, opts)
    got, res = run_ours(fresh_ctx, hdr, [synth.pack_bits(f) for f in frames], opts, "C4.MLV")
    assert res[0].is_dual_iso == 1 and res[0].black_level == 8192
    rec = parity_record("C4 5760x3240 dual-iso amaze-edge alias-map really-bad-pix frame 0", got[0], want[0], 1)
    assert rec["max_abs_diff_dn"] <= 1


def test_c5_geometry_lj92(fresh_ctx, tmp_path):
    w, h, n = 3840, 2160, 2
    hdr = F.make_frame_headers(w, h, video_class=F.VIDEO_CLASS_RAW | F.VIDEO_CLASS_FLAG_LJ92)
    frames = [synth.make_frame(w, h, i, hot_cold=True) for i in range(n)]
    payloads = [synth.lj92_payload(f) for f in frames]
    synth.write_mlv(str(tmp_path / "C5.MLV"), (p.tobytes() for p in payloads), hdr)
    want = reference_frames(tmp_path, "C5.MLV", n, {})
    got, _ = run_ours(fresh_ctx, hdr, payloads, {}, "C5.MLV")
    for i in range(n):
        assert np.array_equal(want[i], frames[i])
        rec = parity_record(f"C5 3840x2160 LJ92 frame {i}", got[i], want[i], 0)
        assert rec["max_abs_diff_dn"] == 0


def test_c2_geometry_reference_process_frame(fresh_ctx, tmp_path):
    """The headline config against the reference's own process_frame (the other C2 tests use the oracle restatement)."""
    w, h, n = 1920, 1080, 3
    opts = dict(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=1)
    hdr, frames = synth.make_clip(str(tmp_path / "C2.MLV"), w, h, n, variant=dict(hot_cold=True, stripes=True))
    want = reference_frames(tmp_path, "C2.MLV", n, opts)
    got, _ = run_ours(fresh_ctx, hdr, [synth.pack_bits(f) for f in frames], opts, "C2.MLV")
    for i in range(n):
        rec = parity_record(f"C2 1920x1080 stripes bad-pix cs3x3 frame {i}", got[i], want[i], 0)
        assert rec["max_abs_diff_dn"] == 0
