"""GPU: the host frame-request path (frame cache -> process_frame -> libmlvfs_b200) through the
mlvb_frames driver, against the oracle.  Mirrors how the FUSE read handler asks for frames."""
import json
import os
import subprocess

import numpy as np
import pytest

from mlvfs_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "mlvfs_b200", "mlvb_frames")


def _run(args, cwd, env=None):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mlvfs_b200", "host")])
    out = subprocess.check_output([EXE] + args, cwd=cwd, text=True, env=dict(os.environ, **(env or {})))
    return json.loads(out.strip().splitlines()[-1])


@pytest.mark.parametrize("readers,prefetch", [(1, 0), (3, 4)])
def test_frame_server_matches_oracle(oracle, tmp_path, readers, prefetch):
    w, h, n = 1920, 1080, 10
    hdr, frames = synth.make_clip(str(tmp_path / "C2.MLV"), w, h, n, variant=dict(hot_cold=True, stripes=True))
    ri = hdr.rawi_hdr.raw_info
    want, _ = oracle.single_iso_chain(frames, ri.black_level, ri.white_level, ri.frame_size,
                                      chroma_smooth_method=3, fix_bad_pixels=1, fix_stripes=1)
    dump = str(tmp_path / "out.raw")
    res = _run([str(tmp_path), "C2.MLV", "--cs3x3", "--bad-pix", "--stripes", f"--readers={readers}",
                f"--prefetch={prefetch}", f"--dump={dump}"], cwd=str(tmp_path))
    assert res["frames"] == n and res["failed"] == 0
    got = np.fromfile(dump, dtype=np.uint16).reshape(n, h, w)
    for i in range(n):
        assert np.array_equal(got[i], want[i]), f"frame {i}"
    if prefetch:
        assert res["prefetch_built"] > 0


def test_frame_server_plain_unpack(tmp_path):
    w, h, n = 640, 360, 5
    hdr, frames = synth.make_clip(str(tmp_path / "C1.MLV"), w, h, n)
    dump = str(tmp_path / "out.raw")
    res = _run([str(tmp_path), "C1.MLV", f"--dump={dump}"], cwd=str(tmp_path))
    got = np.fromfile(dump, dtype=np.uint16).reshape(n, h, w)
    assert res["failed"] == 0 and np.array_equal(got, np.stack(frames))


@pytest.mark.parametrize("gpus,readers,prefetch,batch", [(1, 1, 16, 8), (3, 4, 24, 8), (2, 2, 6, 3)])
def test_multi_context_batched_prefetch_matches_reference(tmp_path, gpus, readers, prefetch, batch):
    """The north-star host shape: one context per GPU (here several contexts on the box's GPU, $MLVB_SHARE_DEVICES),
    frames dealt in chunks, look-ahead chunks built as device batches (the wide fused kernel behind the frame-request
    API), per-clip state primed from frame 0 on every context.  Pixels AND the 64 KiB DNG headers equal the unmodified
    reference's process_frame output for every frame, whatever the reader / prefetch interleaving."""
    from conftest import reference_frames
    w, h, n = 1920, 1080, 40
    hdr, frames = synth.make_clip(str(tmp_path / "C2.MLV"), w, h, n, variant=dict(hot_cold=True, stripes=True))
    want, want_hdr = reference_frames(tmp_path, "C2.MLV", n, dict(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=1), with_headers=True)
    dump, dump_h = str(tmp_path / "out.raw"), str(tmp_path / "hdr.bin")
    res = _run([str(tmp_path), "C2.MLV", "--cs3x3", "--bad-pix", "--stripes", f"--readers={readers}", f"--prefetch={prefetch}",
                f"--gpus={gpus}", f"--batch={batch}", f"--dump={dump}", f"--dump-headers={dump_h}"], cwd=str(tmp_path),
               env={"MLVB_SHARE_DEVICES": "1"})
    assert res["frames"] == n and res["failed"] == 0 and res["gpus"] == gpus
    assert res["prefetch_batches"] > 0 and res["device_batches"] > 0
    got = np.fromfile(dump, dtype=np.uint16).reshape(n, h, w)
    got_hdr = np.fromfile(dump_h, dtype=np.uint8).reshape(n, 65536)
    for i in range(n):
        assert np.array_equal(got[i], want[i]), f"frame {i}"
        assert np.array_equal(got_hdr[i], want_hdr[i]), f"header {i}"


def test_dual_iso_clip_through_the_frame_server(tmp_path):
    """--dual-iso frames are not batchable (per-frame statistics): the look-ahead chunk is pipelined frame by frame.
    Converted frames carry black / white x4 in their DNG headers (main.c:961-965)."""
    from conftest import reference_frames, parity_record
    w, h, n = 960, 540, 9
    hdr, frames = synth.make_clip(str(tmp_path / "D.MLV"), w, h, n, variant=dict(dual_iso=True))
    opts = dict(dual_iso=2, hdr_interpolation_method=1, chroma_smooth=2)
    want, want_hdr = reference_frames(tmp_path, "D.MLV", n, opts, with_headers=True)
    dump, dump_h = str(tmp_path / "out.raw"), str(tmp_path / "hdr.bin")
    res = _run([str(tmp_path), "D.MLV", "--dual-iso", "--mean23", "--cs2x2", "--readers=2", "--prefetch=4", "--gpus=2", "--batch=4",
                f"--dump={dump}", f"--dump-headers={dump_h}"], cwd=str(tmp_path), env={"MLVB_SHARE_DEVICES": "1"})
    assert res["failed"] == 0
    got = np.fromfile(dump, dtype=np.uint16).reshape(n, h, w)
    got_hdr = np.fromfile(dump_h, dtype=np.uint8).reshape(n, 65536)
    for i in range(n):
        assert parity_record(f"frame server dual-iso mean23 cs2x2 {w}x{h} frame {i}", got[i], want[i], 1)["max_abs_diff_dn"] <= 1
        assert np.array_equal(got_hdr[i], want_hdr[i]), f"header {i}"


@pytest.mark.parametrize("codec", ["lj92", "lzma"])
def test_compressed_clips_through_the_frame_server(tmp_path, codec):
    from mlvfs_b200 import mlvformat as F
    w, h, n = 1280, 720, 12
    vc = F.VIDEO_CLASS_RAW | (F.VIDEO_CLASS_FLAG_LJ92 if codec == "lj92" else F.VIDEO_CLASS_FLAG_LZMA)
    hdr = F.make_frame_headers(w, h, video_class=vc)
    frames = [synth.make_frame(w, h, i) for i in range(n)]
    enc = synth.lj92_payload if codec == "lj92" else synth.lzma_payload
    synth.write_mlv(str(tmp_path / "Z.MLV"), (enc(f).tobytes() for f in frames), hdr)
    dump = str(tmp_path / "out.raw")
    res = _run([str(tmp_path), "Z.MLV", "--readers=2", "--prefetch=8", "--batch=4", f"--dump={dump}"], cwd=str(tmp_path))
    assert res["failed"] == 0
    got = np.fromfile(dump, dtype=np.uint16).reshape(n, h, w)
    assert np.array_equal(got, np.stack(frames))
    if codec == "lj92":
        assert res["device_batches"] > 0
