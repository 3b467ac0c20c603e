#!/usr/bin/env python3
"""Decode a few LJ92 frames through the device-batch ABI (for ncu captures).
usage: prof_lj92.py [nframes] [w] [h]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = int(sys.argv[2]) if len(sys.argv) > 2 else 3840
h = int(sys.argv[3]) if len(sys.argv) > 3 else 2160
hdr = F.make_frame_headers(w, h, video_class=F.VIDEO_CLASS_RAW | F.VIDEO_CLASS_FLAG_LJ92)
p = synth.lj92_payload(synth.make_frame(w, h, 0))
stride = (p.size + 1024 + 15) // 16 * 16
packed = np.zeros((n, stride), np.uint8)
packed[:, :p.size] = p
d_in = torch.from_numpy(packed).cuda()
d_out = torch.empty((n, w * h), dtype=torch.int16, device="cuda")
ctx = M.Context(device=0, slots=2)
for _ in range(3):
    ctx.process_batch_device(hdr, M.Options(), "prof.MLV", d_in.data_ptr(), stride, stride, d_out.data_ptr(), w * h, n,
                             torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
ctx.close()
