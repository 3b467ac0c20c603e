#!/usr/bin/env python3
"""Two host threads pushing dual-ISO chunks through mlvb_process_frames with MLVB_TRACE=1: who waits for whom.
usage: MLVB_TRACE=1 python tools/e2e_trace.py [C3|C4] [chunk] [threads] [calls per thread]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import mlvfs_b200 as M
from mlvfs_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nthr = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ncalls = int(sys.argv[4]) if len(sys.argv) > 4 else 3
wl = bench.WORKLOADS[name]
w, h = wl["w"], wl["h"]
npix = w * h
hdr = bench.headers_for(wl)
opts = M.Options(**wl["opts"])
frame = synth.make_frame(w, h, 0, **wl["variant"])
packed = synth.pack_bits(frame).view(np.uint8)
ctx = M.Context(device=0, slots=int(os.environ.get("SLOTS", "8")))
pin_in = M.PinnedBuffer(packed.size)
pin_in.array[:] = packed
pin_out = [M.PinnedBuffer(chunk * npix * 2) for _ in range(nthr)]
hdrs = [hdr] * chunk

def call(t):
    return ctx.process_frames(hdrs, [pin_in.ptr] * chunk, [packed.size] * chunk, opts, "trace.MLV",
                              [pin_out[t].ptr + k * npix * 2 for k in range(chunk)])[0]

assert call(0) == 0 and call(0) == 0          # prime + warm buffers
if nthr > 1:
    ts = [threading.Thread(target=call, args=(t,)) for t in range(nthr)]
    [t.start() for t in ts]; [t.join() for t in ts]
print("---- timed", file=sys.stderr)
t0 = time.perf_counter()
def worker(t):
    for _ in range(ncalls):
        call(t)
ts = [threading.Thread(target=worker, args=(t,)) for t in range(nthr)]
[t.start() for t in ts]; [t.join() for t in ts]
dt = time.perf_counter() - t0
print(f"{name}: {nthr} thread(s) x {ncalls} calls x {chunk} frames: {nthr * ncalls * chunk / dt:.1f} frames/s", file=sys.stderr)
