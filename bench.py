#!/usr/bin/env python3
"""bench.py -- DNG frames/s of the MLVFS per-frame raw path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]           our CUDA path (default workload C2)
  python bench.py --impl reference [...]                        the reference's CPU path (oracle/_ref)
  torchrun --nproc-per-node N bench.py --gpus N ...             one rank per GPU, frames sharded by rank
  python bench.py --workload C1|C2|C3|C4|C5 ...                    the other BASELINE.json configs (not the headline line)

Headline workload (config.workload): BASELINE.json configs[1] = C2 -- 1920x1080 14-bit uncompressed MLV frames
with --stripes --bad-pix --cs3x3 (the full single-ISO correction chain), synthetic input (mlvfs_b200/synth.py).
A "step" is one pass of the hot path over a batch of `frames_per_step` frames.

  value  frames/s with the payloads already resident in HBM (mlvb_process_batch_device), CUDA-event timed
         on the launching stream, max over ranks.
  e2e    frames/s through the host-buffer C ABI (mlvb_submit / mlvb_wait): pinned host payload -> H2D ->
         kernels -> D2H of the finished 16-bit frame, all inside the timed region.
  roofline      the dominant stage of the workload: algorithmic bytes per launch / its
                CUDA-event duration measured in the timed region, vs MEASURED_PEAKS.json.
  cpu_baseline  the unmodified reference (oracle/_ref, process_frame) on this box's host cores on a
                bounded sample of the same workload (falls back to the oracle port if _ref is absent).

The oracle is used here only as the CPU baseline / reference arm, never on the measured GPU path.
"""
import argparse
import collections
import ctypes as C
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> geometry, options (fields of struct mlvfs), input variant, payload codec, algorithmic bytes per
# pixel of the whole chain and of the dominant stage
WORKLOADS = {
    "C1": dict(w=1920, h=1080, opts={}, variant={}, codec="raw", chain_bpp=3.75, stage="unpack", stage_bpp=3.75,
               desc="C1: 1920x1080 14-bit uncompressed MLV, plain unpack -> DNG", frames=256,
               kernel="unpack_groups_kernel<14>"),
    "C2": dict(w=1920, h=1080, opts=dict(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=1),
               variant=dict(hot_cold=True, stripes=True), codec="raw", chain_bpp=3.75, stage="chroma", stage_bpp=3.75,
               desc="C2: 1920x1080 14-bit uncompressed MLV, --stripes --bad-pix --cs3x3", frames=256,
               kernel="fused3_wide_kernel (unpack + bad-pixel patches + 3x3 median chroma smoothing + stripes, one pass; persistent, EV tables in shared memory)"),
    "C3": dict(w=3840, h=1536, opts=dict(dual_iso=2, hdr_interpolation_method=1, chroma_smooth=5),
               variant=dict(dual_iso=True), codec="raw", chain_bpp=7.25, stage="dualiso", stage_bpp=7.25,
               desc="C3: 3840x1536 14-bit dual-ISO MLV, --dual-iso --mean23 --cs5x5 (alias map on)", frames=8,
               kernel="dual-ISO stage (statistics + mean23 + 2x cs5x5 on 20-bit planes + alias map + blend)"),
    "C4": dict(w=5760, h=3240, opts=dict(dual_iso=2, hdr_interpolation_method=0, fix_bad_pixels=2),
               variant=dict(dual_iso=True, hot_cold=True), codec="raw", chain_bpp=7.25, stage="dualiso", stage_bpp=7.25,
               desc="C4: 5760x3240 14-bit dual-ISO MLV, --dual-iso --amaze-edge --alias-map --really-bad-pix", frames=4,
               kernel="dual-ISO stage (statistics + AMaZE + edge-directed interpolation + alias map + blend)"),
    "C5": dict(w=3840, h=2160, opts={}, variant={}, codec="lj92", chain_bpp=2.9, stage="lj92", stage_bpp=2.9,
               desc="C5: 3840x2160 LJ92-compressed MLV, plain decode -> DNG", frames=64,
               kernel="LJ92 stage (unstuff + self-synchronising parallel Huffman decode + wavefront prediction + untile, 14 launches)"),
}
METRIC = "DNG frames/sec per B200 and at 1/2/4/8 GPUs; achieved HBM GB/s vs peak"   # BASELINE.json metric


def measured_traffic(workload, frames):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f).get(workload)
        return float(t["dram_bytes_per_frame"]) * frames if t else None
    except Exception:
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU during the timed region (NVML)."""

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.002)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_frames(wl, n):
    from mlvfs_b200 import synth
    return [synth.make_frame(wl["w"], wl["h"], i, **wl["variant"]) for i in range(n)]


def make_payloads(wl, frames):
    """VIDF payloads as uint8 rows of a common 16-byte aligned stride (+1 KiB tail for LJ92 staging)."""
    from mlvfs_b200 import synth
    if wl["codec"] == "lj92":
        raw = [synth.lj92_payload(f) for f in frames]
    else:
        raw = [synth.pack_bits(f).view(np.uint8) for f in frames]
    stride = (max(p.size for p in raw) + (1024 if wl["codec"] == "lj92" else 0) + 15) // 16 * 16
    out = np.zeros((len(raw), stride), np.uint8)
    for i, p in enumerate(raw):
        out[i, :p.size] = p
    return out, [int(p.size) for p in raw]


def headers_for(wl):
    from mlvfs_b200 import mlvformat as F
    vc = F.VIDEO_CLASS_RAW | (F.VIDEO_CLASS_FLAG_LJ92 if wl["codec"] == "lj92" else 0)
    return F.make_frame_headers(wl["w"], wl["h"], video_class=vc)


# ----------------------------------------------------------------------------------------------
# reference / CPU arm

def reference_runner(wl):
    """Returns (kind, threads, fn(nframes) -> seconds) timing the reference CPU path on this workload."""
    from mlvfs_b200 import synth
    from oracle import pyoracle as O
    ref = O.load_ref()
    cores = os.cpu_count() or 1
    hdr = headers_for(wl)
    ri = hdr.rawi_hdr.raw_info
    npix = wl["w"] * wl["h"]
    nclip = 4 if npix > 4e6 else 8
    frames = make_frames(wl, nclip)
    o = wl["opts"]
    if ref is not None:
        threads = min(cores, 32)
        tmp = tempfile.mkdtemp(prefix="mlvb_ref_")
        payloads, sizes = make_payloads(wl, frames)
        synth.write_mlv(os.path.join(tmp, "W.MLV"), (payloads[i, :sizes[i]].tobytes() for i in range(nclip)), hdr)
        ref.ref_set_mlv_dir(tmp.encode())
        ref.ref_set_options(o.get("chroma_smooth", 0), o.get("fix_bad_pixels", 0), o.get("fix_stripes", 0), o.get("dual_iso", 0),
                            o.get("hdr_interpolation_method", 0), o.get("hdr_no_fullres", 0), o.get("hdr_no_alias_map", 0),
                            o.get("fix_pattern_noise", 0), o.get("deflicker", 0))
        bufs = [np.empty(npix, np.uint16) for _ in range(threads)]

        def one(tid, idx):
            ref.ref_process_frame(b"/W.MLV/W_%06d.dng" % (idx % nclip), bufs[tid].ctypes.data_as(C.c_void_p),
                                  bufs[tid].nbytes, None)

        with O.quiet_stdout():
            one(0, 0)               # frame 0 first: creates the per-clip state on one thread

        def run(nframes):
            counter = iter(range(nframes))
            lock = threading.Lock()

            def worker(tid):
                while True:
                    with lock:
                        i = next(counter, None)
                    if i is None:
                        return
                    one(tid, i)

            with O.quiet_stdout():
                ts = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
                t0 = time.perf_counter()
                [t.start() for t in ts]
                [t.join() for t in ts]
                return time.perf_counter() - t0

        return "reference", threads, run

    # oracle port (no oracle/_ref on this machine): single-ISO chains only
    threads = min(cores, 16)
    state = {}
    kw = dict(chroma_smooth_method=o.get("chroma_smooth", 0), fix_bad_pixels=o.get("fix_bad_pixels", 0),
              fix_stripes=o.get("fix_stripes", 0))
    if o.get("dual_iso") or wl["codec"] != "raw":
        raise SystemExit("bench.py: oracle/_ref is required for the CPU arm of this workload")
    O.single_iso_chain(frames[:1], ri.black_level, ri.white_level, ri.frame_size, state=state, **kw)

    def run(nframes):
        counter = iter(range(nframes))
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    i = next(counter, None)
                if i is None:
                    return
                O.single_iso_chain([frames[i % nclip]], ri.black_level, ri.white_level, ri.frame_size, state=state, **kw)

        ts = [threading.Thread(target=worker) for _ in range(threads)]
        t0 = time.perf_counter()
        [t.start() for t in ts]
        [t.join() for t in ts]
        return time.perf_counter() - t0

    return "port", threads, run


def cpu_baseline(wl, budget_s=12.0):
    kind, threads, run = reference_runner(wl)
    n = max(threads, 4)
    dt = run(n)                                   # calibration pass doubles as warm-up
    total = int(min(max(n, n * budget_s / max(dt, 1e-3)), 50 * threads))
    dt = run(total)
    what = "oracle/_ref process_frame (unmodified reference, gcc -O2)" if kind == "reference" else "the oracle port"
    return {"value": total / dt, "unit": "frames/s", "cores": threads, "kind": kind,
            "sample": f"{total} frames of {wl['desc'].split(':')[0]} through {what} on {threads} threads, {dt:.1f} s"}


def run_reference(args, wl):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    kind, threads, run = reference_runner(wl)
    per_step = max(threads * 2, 8)
    run(per_step)
    t = 0.0
    for _ in range(args.steps):
        t += run(per_step)
    fps = per_step * args.steps / t
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": wl["desc"], "frames_per_step": per_step},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"{per_step} frames/step x {args.steps} steps on {threads} host threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------
# our arm

def run_ours(args, wl):
    import torch
    import mlvfs_b200 as M

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- mlvfs_b200 has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w, h = wl["w"], wl["h"]
    npix = w * h
    B = args.frames_per_step or wl["frames"]
    hdr = headers_for(wl)
    opts = M.Options(**wl["opts"])
    distinct = min(B, 4 if npix > 4e6 else 8)
    base, sizes = make_payloads(wl, make_frames(wl, distinct))
    stride = base.shape[1]
    packed = np.ascontiguousarray(base[np.arange(B) % distinct])
    in_bytes = float(np.mean(sizes))
    ctx = M.Context(device=local, slots=args.slots)
    d_in = torch.from_numpy(packed).cuda()
    d_out = torch.empty((B, npix), dtype=torch.int16, device="cuda")
    stream = torch.cuda.Stream()
    clip = "bench_%s.MLV" % args.workload

    def step():
        ctx.process_batch_device(hdr, opts, clip, d_in.data_ptr(), stride, stride, d_out.data_ptr(), npix, B, stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    W_ = max(args.warmup, 3)
    for _ in range(W_):
        step()
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count()
    ctx.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    stages = ctx.profile_end()
    launches = ctx.launch_count() - launches0
    clocks = sampler.finish()
    t = torch.tensor([ms], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * args.steps * world / (ms * 1e-3)

    # ---- e2e through the host-buffer ABI: pinned payloads in, finished frames out
    pin_in = M.PinnedBuffer(B * stride)
    pin_in.array[:] = packed.reshape(-1)
    depth = args.slots
    pin_out = [M.PinnedBuffer(npix * 2) for _ in range(depth)]

    def e2e_step():
        q = collections.deque()
        for f in range(B):
            if len(q) == depth:
                ctx.wait(q.popleft())
            tk = ctx.submit(hdr, C.c_void_p(pin_in.ptr + f * stride), stride, opts, clip, C.c_void_p(pin_out[f % depth].ptr))
            if tk < 0:
                raise RuntimeError(f"mlvb_submit failed: {tk}")
            q.append(tk)
        while q:
            ctx.wait(q.popleft())

    light = wl["codec"] == "raw" and not wl["opts"].get("dual_iso")
    e2e_steps = max(1, min(args.steps, int(np.ceil(2000 / B)))) if light else 1
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda")
    if dist is not None:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = B * e2e_steps * world / float(t.item())
    checksum = int(pin_out[0].array.view(np.uint16)[::4099].astype(np.uint64).sum())

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        chain_bytes = wl["chain_bpp"] * npix if wl["codec"] == "raw" else in_bytes + 2 * npix
        roof = None
        if wl["stage"] in stages:
            tot_ms, spans = stages[wl["stage"]]
            per_launch_s = tot_ms * 1e-3 / args.steps          # one step = one batch through this stage
            stage_bytes = (wl["stage_bpp"] * npix if wl["codec"] == "raw" else in_bytes + 2 * npix) * B
            achieved = stage_bytes / per_launch_s / 1e9
            roof = {"bound": "hbm", "kernel": wl["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": measured_traffic(args.workload, B), "peak_source": peak_src, "launch_ms": per_launch_s * 1e3,
                    "algorithmic_bytes_per_launch": stage_bytes,
                    "stage_ms_per_step": {k: v[0] / args.steps for k, v in stages.items()},
                    "chain_achieved_gbs": chain_bytes * value / world / 1e9,
                    "chain_frac": chain_bytes * value / world / 1e9 / peak}
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": W_,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16 pixels / int32 EV-LUT arithmetic" + (" / fp64 blends" if wl["opts"].get("dual_iso") else ""),
            "data": "synthetic",
            "config": {"workload": wl["desc"], "frames_per_step": B, "sharding": f"frames by rank, {world} rank(s), no collective",
                       "cache": f"inputs+outputs per step {B * chain_bytes / 1e6:.0f} MB "
                                + ("> 126 MB L2 (no flush needed)" if B * chain_bytes > 252e6 else "(fits L2: treat as warm-cache)")},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(B * in_bytes), "d2h_bytes_per_step": B * npix * 2,
                    "frames_in_flight": depth, "steps": e2e_steps, "checksum": checksum},
            "gpu_launches": launches,
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line))
    for p in pin_out:
        p.free()
    pin_in.free()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames-per-step", type=int, default=0)
    ap.add_argument("--slots", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
