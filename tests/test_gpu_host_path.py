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


def _run(args, cwd):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mlvfs_b200", "host")])
    out = subprocess.check_output([EXE] + args, cwd=cwd, text=True)
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
