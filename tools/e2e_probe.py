"""Probe: end-to-end frames/s of the host-buffer ABI for several option sets (same frames, pinned buffers)."""
import collections, ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F, synth

w, h, B, depth = 1920, 1080, 128, 8
hdr = F.make_frame_headers(w, h)
frames = [synth.make_frame(w, h, i % 4, hot_cold=True, stripes=True) for i in range(4)]
packed = np.stack([synth.pack_bits(frames[i % 4]) for i in range(B)])
stride = packed.shape[1] * 2
pin_in = M.PinnedBuffer(B * stride)
pin_in.array[:] = packed.view(np.uint8).reshape(-1)
pin_out = [M.PinnedBuffer(w * h * 2) for _ in range(depth)]
ctx = M.Context(device=0, slots=depth)
for name, kw in [("plain", {}), ("cs3", dict(chroma_smooth=3)), ("stripes", dict(fix_stripes=1)), ("badpix", dict(fix_bad_pixels=1)),
                 ("c2", dict(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=1)), ("plain", {})]:
    opts = M.Options(**kw)
    def run():
        q = collections.deque()
        for f in range(B):
            if len(q) == depth:
                ctx.wait(q.popleft())
            q.append(ctx.submit(hdr, C.c_void_p(pin_in.ptr + f * stride), stride, opts, "probe_" + name, C.c_void_p(pin_out[f % depth].ptr)))
        while q:
            ctx.wait(q.popleft())
    run(); run()
    t0 = time.perf_counter()
    for _ in range(6):
        run()
    dt = time.perf_counter() - t0
    print(f"{name:8s} {6 * B / dt:9.0f} frames/s")
