#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: run the unmodified reference's process_frame (oracle/_ref, main.c:908-1005) on an MLV clip in
a FRESH process and dump the frames.  A fresh process matters: the reference keeps process-lifetime state (the first
dual-ISO frame's white level baked into its 20-bit tables hdr.c:1089-1093, bad-pixel ring, stripes list, rand()).

    python tests/refrun.py <mlv_dir> <clip> <nframes> <out.npy> cs badpix stripes dual_iso interp no_fullres no_alias pattern deflicker
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import pyoracle as O
    mlv_dir, clip, n, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    opts = [int(v) for v in sys.argv[5:14]]
    ref = O.load_ref()
    if ref is None:
        print("oracle/_ref is not built", file=sys.stderr)
        return 3
    ref.ref_set_mlv_dir(mlv_dir.encode())
    ref.ref_set_options(*opts)
    hdr = (C.c_uint8 * 4096)()
    stem = os.path.splitext(clip)[0]
    from mlvfs_b200 import mlvformat as F
    fh = F.FrameHeaders()
    assert ref.ref_get_frame_headers(os.path.join(mlv_dir, clip).encode(), 0, C.byref(fh))
    w, h = fh.rawi_hdr.xRes, fh.rawi_hdr.yRes
    frames = np.zeros((n, h, w), np.uint16)
    headers = np.zeros((n, 65536), np.uint8)
    with O.quiet_stdout():
        for i in range(n):
            got = ref.ref_process_frame(f"/{clip}/{stem}_{i:06d}.dng".encode(), frames[i].ctypes.data_as(C.c_void_p),
                                        frames[i].nbytes, headers[i].ctypes.data_as(C.c_void_p))
            if got != frames[i].nbytes:
                print(f"reference process_frame failed on frame {i}", file=sys.stderr)
                return 4
    np.save(out, frames)
    np.save(out + ".headers.npy", headers)
    return 0


if __name__ == "__main__":
    sys.exit(main())
