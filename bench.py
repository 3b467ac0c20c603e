#!/usr/bin/env python3
"""bench.py -- DNG frames/s of the MLVFS per-frame raw path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]           our CUDA path: headline workload C2 + every other config
  python bench.py --impl reference [...]                        the reference's CPU path (oracle/_ref), same workloads
  torchrun --nproc-per-node N bench.py --gpus N ...             one rank per GPU, frames sharded by rank
  python bench.py --workload C1|C2|C3|C4|C5 --only ...           one config as the headline line, nothing else

Headline workload (config.workload): BASELINE.json configs[1] = C2 -- 1920x1080 14-bit uncompressed MLV frames
with --stripes --bad-pix --cs3x3 (the full single-ISO correction chain), synthetic input (mlvfs_b200/synth.py).
A "step" is one pass of the hot path over a batch of `frames_per_step` frames.

  value      frames/s with the payloads already resident in HBM (mlvb_process_batch_device), CUDA-event timed
             on the launching stream, max over ranks.  `sustained` = the same step repeated for >= 1 s.
  e2e        frames/s through the host-buffer C ABI (mlvb_process_frames, what the --prefetch queue calls): pinned
             host payloads -> H2D -> kernels -> D2H of the finished 16-bit frames, all inside the timed region.
  roofline   the dominant stage of the workload: algorithmic bytes per launch / its CUDA-event
             duration measured in the timed region, vs MEASURED_PEAKS.json.
  workloads  the other BASELINE.json configs (C1, C3, C4, C5), each with value / e2e / roofline, at every N.
  host_path  frames/s of the frame-request path itself (mlvb_frames: frame cache -> process_frame -> per-GPU
             contexts, batched prefetch) in ONE process using all N GPUs, run by rank 0.
  cpu_baseline  the unmodified reference (oracle/_ref, process_frame) on this box's host cores on a
                bounded sample of the headline workload (falls back to the oracle port if _ref is absent).

The oracle is used here only as the CPU baseline / reference arm, never on the measured GPU path.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
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
               desc="C1: 1920x1080 14-bit uncompressed MLV, plain unpack -> DNG", frames=256, e2e_chunk=8,
               kernel="unpack_groups_kernel<14>", cli=[]),
    "C2": dict(w=1920, h=1080, opts=dict(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=1),
               variant=dict(hot_cold=True, stripes=True), codec="raw", chain_bpp=3.75, stage="chroma", stage_bpp=3.75,
               desc="C2: 1920x1080 14-bit uncompressed MLV, --stripes --bad-pix --cs3x3", frames=512, e2e_chunk=8,
               kernel="fused3_wide_kernel (unpack + bad-pixel patches + 3x3 median chroma smoothing + stripes, one pass; persistent, EV tables in shared memory)",
               cli=["--cs3x3", "--bad-pix", "--stripes"]),
    "C3": dict(w=3840, h=1536, opts=dict(dual_iso=2, hdr_interpolation_method=1, chroma_smooth=5),
               variant=dict(dual_iso=True), codec="raw", chain_bpp=7.25, stage="dualiso", stage_bpp=7.25,
               desc="C3: 3840x1536 14-bit dual-ISO MLV, --dual-iso --mean23 --cs5x5 (alias map on)", frames=16, e2e_chunk=16,
               kernel="dual-ISO stage (statistics + mean23 + 2x cs5x5 on 20-bit planes + alias map + blend)",
               cli=["--dual-iso", "--mean23", "--cs5x5"]),
    "C4": dict(w=5760, h=3240, opts=dict(dual_iso=2, hdr_interpolation_method=0, fix_bad_pixels=2),
               variant=dict(dual_iso=True, hot_cold=True), codec="raw", chain_bpp=7.25, stage="dualiso", stage_bpp=7.25,
               desc="C4: 5760x3240 14-bit dual-ISO MLV, --dual-iso --amaze-edge --alias-map --really-bad-pix", frames=8, e2e_chunk=8,
               kernel="dual-ISO stage (statistics + AMaZE + edge-directed interpolation + alias map + blend)",
               cli=["--dual-iso", "--amaze-edge", "--alias-map", "--really-bad-pix"]),
    "C5": dict(w=3840, h=2160, opts={}, variant={}, codec="lj92", chain_bpp=2.9, stage="lj92", stage_bpp=2.9,
               desc="C5: 3840x2160 LJ92-compressed MLV, plain decode -> DNG", frames=256, e2e_chunk=16,
               kernel="LJ92 stage (unstuff + self-synchronising parallel Huffman decode + separated predictor-6 scans + untile)",
               cli=[]),
}
METRIC = "DNG frames/sec per B200 and at 1/2/4/8 GPUs; achieved HBM GB/s vs peak"   # BASELINE.json metric
DTYPE = "u16 pixels / int32 EV-LUT arithmetic (dual-ISO configs: + fp64 blends, fp32 AMaZE)"


def config_of(wl, frames_per_step, world):
    """The `config` object: the same keys and values in both arms (the driver compares them)."""
    return {"workload": wl["desc"], "frames_per_step": frames_per_step,
            "sharding": f"frames by rank, {world} rank(s), no collective",
            "cache": "inputs + outputs of a step exceed the 126 MB L2 (no flush needed)"
                     if frames_per_step * wl["chain_bpp"] * wl["w"] * wl["h"] > 252e6 else "inputs + outputs of a step fit L2 partly: warm-cache"}


def measured_traffic(workload, frames):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), with its
    source -- a profile constant, not measured in this run -- or (None, None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f).get(workload)
        return (float(t["dram_bytes_per_frame"]) * frames, t["source"]) if t else (None, None)
    except Exception:
        return None, None


def profile_fractions(workload):
    """Issue / ALU-pipe fractions of the dominant kernel from the committed ncu summaries (profiles/ncu_fractions.json),
    for the stages that are not HBM-bound; constants with their source, not measured in this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_fractions.json")) as f:
            return json.load(f).get(workload)
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


_PAYLOADS = {}


def distinct_payloads(wl):
    """The workload's distinct synthetic payloads (generated once per process: 18 MP frames take seconds in numpy)."""
    key = wl["desc"]
    if key not in _PAYLOADS:
        _PAYLOADS[key] = make_payloads(wl, make_frames(wl, distinct_frames(wl)))
    return _PAYLOADS[key]


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


def distinct_frames(wl):
    return 4 if wl["w"] * wl["h"] > 4e6 else 8


def write_clip(wl, path, nframes):
    """A synthetic MLV of `nframes` VIDF blocks cycling through the workload's distinct frames."""
    from mlvfs_b200 import synth
    nd = distinct_frames(wl)
    payloads, sizes = distinct_payloads(wl)
    blobs = [payloads[i, :sizes[i]].tobytes() for i in range(nd)]
    synth.write_mlv(path, (blobs[i % nd] for i in range(nframes)), headers_for(wl))


# ----------------------------------------------------------------------------------------------
# reference / CPU arm

def reference_runner(wl):
    """Returns (kind, threads, fn(nframes) -> seconds) timing the reference CPU path on this workload."""
    from oracle import pyoracle as O
    ref = O.load_ref()
    cores = os.cpu_count() or 1
    hdr = headers_for(wl)
    ri = hdr.rawi_hdr.raw_info
    npix = wl["w"] * wl["h"]
    nclip = distinct_frames(wl)
    o = wl["opts"]
    if ref is not None:
        threads = min(cores, 32)
        tmp = tempfile.mkdtemp(prefix="mlvb_ref_")
        write_clip(wl, os.path.join(tmp, "W.MLV"), nclip)
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
    frames = make_frames(wl, nclip)
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


def cpu_sample(wl, budget_s):
    """Bounded sample of the reference CPU path: about budget_s seconds of frames on all host threads."""
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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kind, threads, run = reference_runner(wl)
    per_step = max(threads * 2, 8)
    for _ in range(max(min(args.warmup, 1), 1)):
        run(per_step)
    t = 0.0
    for _ in range(args.steps):
        t += run(per_step)
    fps = per_step * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": config_of(wl, args.frames_per_step or wl["frames"], world),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"{per_step} frames/step x {args.steps} steps on {threads} host threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.only:
        others = {}
        for name in sorted(WORKLOADS):
            if name == args.workload:
                continue
            heavy = bool(WORKLOADS[name]["opts"].get("dual_iso"))
            others[name] = dict(cpu_sample(WORKLOADS[name], 12.0 if heavy else 5.0), config={"workload": WORKLOADS[name]["desc"]})
        line["workloads"] = others
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm

class Rig:
    """Per-process state of our arm: device, optional process group, helpers."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- mlvfs_b200 has no CPU path")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure(rig, M, name, wl, steps, warmup, slots, frames_per_step=0, sustain_s=1.0, e2e_s=1.5, sample_clocks=False,
            e2e_threads=3, e2e_chunk=0, quick=False):
    """One workload on this rank's GPU: device-resident steps (timed + sustained), per-stage times, e2e."""
    torch = rig.torch
    w, h = wl["w"], wl["h"]
    npix = w * h
    B = frames_per_step or wl["frames"]
    hdr = headers_for(wl)
    opts = M.Options(**wl["opts"])
    distinct = min(B, distinct_frames(wl))
    base, sizes = distinct_payloads(wl)
    stride = base.shape[1]
    # this rank's shard of a clip of B * world frames (frame n -> rank n mod world), primed with the clip's frame 0
    from mlvfs_b200 import sharding
    own = np.array(sharding.frames_for_rank(B * rig.world, rig.rank, rig.world))
    packed = np.ascontiguousarray(base[own % distinct])
    in_bytes = float(np.mean(sizes))
    ctx = M.Context(device=rig.local, slots=slots)
    d_in = torch.from_numpy(packed).cuda()
    d_out = torch.empty((B, npix), dtype=torch.int16, device="cuda")
    stream = torch.cuda.Stream()
    clip = "bench_%s.MLV" % name

    def step():
        ctx.process_batch_device(hdr, opts, clip, d_in.data_ptr(), stride, stride, d_out.data_ptr(), npix, B, stream.cuda_stream)

    d_prime = torch.from_numpy(np.ascontiguousarray(base[np.array(sharding.prime_frames(rig.rank, rig.world)) % distinct])).cuda()
    ctx.process_batch_device(hdr, opts, clip, d_prime.data_ptr(), stride, stride, d_out.data_ptr(), npix, d_prime.shape[0],
                             stream.cuda_stream)                 # per-clip state from frame 0 on every rank
    W_ = max(warmup, 3)
    for _ in range(W_):
        step()
    rig.barrier()

    sampler = ClockSampler(rig.local) if sample_clocks else None
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count()
    ctx.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rig.barrier()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    rig.barrier()
    ms = rig.max_over_ranks(e0.elapsed_time(e1))
    stages = ctx.profile_end()
    launches = ctx.launch_count() - launches0
    value = B * steps * rig.world / (ms * 1e-3)

    if quick:                                             # profiling runs (ncu): the timed steps only
        ctx.close()
        return {"value": value, "unit": "frames/s", "ms_per_step": ms / steps, "frames_per_step": B, "steps": steps,
                "gpu_launches": launches, "stage_ms_per_step": {k: v[0] / steps for k, v in stages.items()}}
    # ---- sustained: the same step back to back for >= sustain_s (clocks and power settle; a 12 ms region cannot show that)
    n_sus = max(steps, int(np.ceil(sustain_s * 1e3 / max(ms / steps, 1e-3))))
    rig.barrier()
    e0.record(stream)
    for _ in range(n_sus):
        step()
    e1.record(stream)
    rig.barrier()
    ms_sus = rig.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.finish() if sampler else None
    sustained = {"value": B * n_sus * rig.world / (ms_sus * 1e-3), "unit": "frames/s", "steps": n_sus, "seconds": ms_sus * 1e-3}

    # ---- e2e through the host-buffer ABI: pinned payloads in, finished frames out, chunks of frames per call
    # single-ISO chunks: 8 frames per call pipeline best on one GPU; with several ranks sharing the host's cores and PCIe
    # uplinks fewer, larger calls do (measured at 8 GPUs: 32 frames per call 20.2 k C2 frames/s, 8 per call 13.1 k)
    chunk = e2e_chunk or wl["e2e_chunk"]
    if not e2e_chunk and rig.world > 1 and not wl["opts"].get("dual_iso"):
        chunk = max(chunk, 32)
    chunk = min(chunk, B)
    nthreads = e2e_threads
    pin_in = M.PinnedBuffer(B * stride)
    pin_in.array[:] = packed.reshape(-1)
    pin_out = [M.PinnedBuffer(chunk * npix * 2) for _ in range(nthreads)]
    hdrs = [hdr] * chunk
    nbytes = [sizes[f % distinct] for f in range(B)]
    nchunks = B // chunk
    # at least two calls in flight (two host threads), so that one call's copies overlap the other's kernels: a pass
    # pushes the step's chunks, twice over if the step is a single chunk
    pass_chunks = max(nchunks, min(2, nthreads))
    errors = []

    def e2e_pass():
        """All B frames once: nthreads host threads, each pushing whole chunks through mlvb_process_frames."""
        nxt = iter(range(pass_chunks))
        lock = threading.Lock()

        def worker(t):
            dsts = [pin_out[t].ptr + k * npix * 2 for k in range(chunk)]
            while True:
                with lock:
                    c = next(nxt, None)
                if c is None:
                    return
                f0 = (c % nchunks) * chunk
                rc, _ = ctx.process_frames(hdrs, [pin_in.ptr + (f0 + k) * stride for k in range(chunk)], nbytes[f0:f0 + chunk],
                                           opts, clip, dsts)
                if rc != 0:
                    errors.append(rc)

        ts = [threading.Thread(target=worker, args=(t,)) for t in range(min(nthreads, pass_chunks))]
        [t.start() for t in ts]
        [t.join() for t in ts]

    e2e_pass()                                            # warm-up: buffers, batch slots
    t0 = time.perf_counter()
    e2e_pass()
    one_pass = time.perf_counter() - t0
    e2e_steps = max(1, int(np.ceil(e2e_s / max(one_pass, 1e-4))))
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_pass()
    torch.cuda.synchronize()
    dt = rig.max_over_ranks(time.perf_counter() - t0)
    if rig.dist is not None:
        rig.dist.barrier()
    if errors:
        raise RuntimeError(f"mlvb_process_frames failed: {errors[:3]}")
    e2e = pass_chunks * chunk * e2e_steps * rig.world / dt
    checksum = int(pin_out[0].array.view(np.uint16)[::4099].astype(np.uint64).sum())

    res = {"value": value, "unit": "frames/s", "ms_per_step": ms / steps, "frames_per_step": B, "steps": steps,
           "sustained": sustained, "gpu_launches": launches,
           "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(sum(nbytes[(c % nchunks) * chunk + k] for c in range(pass_chunks) for k in range(chunk))),
                   "d2h_bytes_per_step": pass_chunks * chunk * npix * 2, "frames_per_step": pass_chunks * chunk, "frames_per_call": chunk,
                   "host_threads": min(nthreads, pass_chunks),
                   "steps": e2e_steps, "seconds": dt, "api": "mlvb_process_frames (pinned host buffers)", "checksum": checksum}}
    if clocks is not None:
        res["clocks"] = clocks
    peak, peak_src = measured_peak_gbs()
    chain_bytes = wl["chain_bpp"] * npix if wl["codec"] == "raw" else in_bytes + 2 * npix
    if wl["stage"] in stages:
        tot_ms, spans = stages[wl["stage"]]
        per_launch_s = tot_ms * 1e-3 / steps              # one step = one batch through this stage
        stage_bytes = (wl["stage_bpp"] * npix if wl["codec"] == "raw" else in_bytes + 2 * npix) * B
        achieved = stage_bytes / per_launch_s / 1e9
        traffic, traffic_src = measured_traffic(name, B)
        res["roofline"] = {"bound": "hbm", "kernel": wl["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                           "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                           "launch_ms": per_launch_s * 1e3, "algorithmic_bytes_per_launch": stage_bytes,
                           "stage_ms_per_step": {k: v[0] / steps for k, v in stages.items()},
                           "chain_achieved_gbs": chain_bytes * value / rig.world / 1e9,
                           "chain_frac": chain_bytes * value / rig.world / 1e9 / peak,
                           "sustained_chain_frac": chain_bytes * sustained["value"] / rig.world / 1e9 / peak}
        fr = profile_fractions(name)
        if fr:
            res["roofline"]["not_hbm_bound"] = fr
    for p in pin_out:
        p.free()
    pin_in.free()
    ctx.close()
    del d_in, d_out
    torch.cuda.empty_cache()
    return res


def host_path(wl, gpus, nframes, prefetch=64, batch=8, readers=8, repeat=3, workers=0):
    """mlvb_frames (the frame-request path: frame cache -> process_frame -> per-GPU contexts with batched prefetch) on a
    synthetic clip, ONE process using `gpus` GPUs.  Returns its JSON report."""
    exe = os.path.join(ROOT, "mlvfs_b200", "mlvb_frames")
    if not os.path.exists(exe):
        return {"unavailable": "mlvfs_b200/mlvb_frames is not built"}
    import shutil
    need = int(1.1 * nframes * wl["w"] * wl["h"] * 1.75) + (64 << 20)
    base = None
    for cand in ("/dev/shm", tempfile.gettempdir()):
        try:
            if os.access(cand, os.W_OK) and shutil.disk_usage(cand).free > need:
                base = cand
                break
        except OSError:
            pass
    if base is None:
        return {"unavailable": f"no scratch directory with {need >> 20} MiB free for the synthetic clip"}
    with tempfile.TemporaryDirectory(prefix="mlvb_host_", dir=base) as d:
        write_clip(wl, os.path.join(d, "H.MLV"), nframes)
        cmd = [exe, d, "H.MLV"] + wl["cli"] + [f"--prefetch={prefetch}", f"--batch={batch}", f"--readers={readers}", f"--gpus={gpus}",
                                              f"--repeat={repeat}", f"--workers={workers}"]
        try:
            out = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=600)
            rep = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception as e:                                 # noqa: BLE001
            return {"unavailable": f"mlvb_frames failed: {e}"}
    return {"value": rep["sustained_fps"], "unit": "frames/s", "first_pass_fps": rep["fps"], "gpus": rep["gpus"], "frames": rep["frames"],
            "passes": rep["passes"], "failed": rep["failed"], "prefetch": prefetch, "batch": batch, "readers": readers,
            "device_batches": rep["device_batches"], "builder_us_per_frame": rep.get("builder_us_per_frame"),
            "reader_copy_us_per_frame": rep.get("reader_copy_us_per_frame"),
            "what": "mlvb_frames: get_or_create_image_buffer -> process_frame / process_frame_batch, one process, clip in page cache"}


def run_ours(args, wl):
    import mlvfs_b200 as M
    rig = Rig()
    if args.quick:
        print(json.dumps(measure(rig, M, args.workload, wl, args.steps, args.warmup, args.slots, args.frames_per_step, quick=True)))
        return
    head = measure(rig, M, args.workload, wl, args.steps, args.warmup, args.slots, args.frames_per_step, sample_clocks=True,
                   e2e_threads=args.e2e_threads, e2e_chunk=args.e2e_chunk)
    others = {}
    if not args.only:
        for name in sorted(WORKLOADS):
            if name == args.workload:
                continue
            o = WORKLOADS[name]
            heavy = bool(o["opts"].get("dual_iso"))
            r = measure(rig, M, name, o, steps=3 if heavy else 10, warmup=3, slots=args.slots, sustain_s=1.0, e2e_s=1.5)
            r["config"] = {"workload": o["desc"], "frames_per_step": r.pop("frames_per_step")}
            others[name] = r
    hp = None
    if not args.only and not args.no_host_path:
        # the frame-request path in ONE process over all of this job's GPUs; the other ranks wait at the barrier
        rig.torch.cuda.empty_cache()
        M.lib().mlvb_host_pool_trim()            # the frame server is another process: hand back this one's pinned pool first
        if rig.rank == 0:
            hp = {args.workload: host_path(wl, rig.world, 256)}
            if rig.world > 1:
                hp[args.workload + "_1gpu"] = host_path(wl, 1, 256)
            # look-ahead deep enough to keep two chunks per GPU in flight
            hp["C4"] = host_path(WORKLOADS["C4"], rig.world, 12 * rig.world, prefetch=8 * rig.world, batch=4, readers=4, repeat=2,
                                 workers=2 * rig.world)
            if rig.world > 1:
                hp["C4_1gpu"] = host_path(WORKLOADS["C4"], 1, 12, prefetch=8, batch=4, readers=4, repeat=2, workers=2)
        rig.barrier()
    if rig.rank == 0:
        B = head.pop("frames_per_step")
        line = {
            "metric": METRIC, "value": head["value"], "unit": "frames/s", "n_gpus": rig.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config_of(wl, B, rig.world),
            "clocks": head.get("clocks"), "sustained": head["sustained"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
            "roofline": head.get("roofline"),
        }
        if others:
            line["workloads"] = others
        if hp:
            line["host_path"] = hp
        if rig.world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sample(wl, 10.0)
        print(json.dumps(line))
    if rig.dist is not None:
        rig.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--only", action="store_true", help="only the headline workload: no other configs, no host path")
    ap.add_argument("--frames-per-step", type=int, default=0)
    ap.add_argument("--slots", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-host-path", action="store_true")
    ap.add_argument("--quick", action="store_true", help="warm-up + timed steps of the headline workload only (for ncu runs)")
    ap.add_argument("--e2e-threads", type=int, default=3)       # measured: 8-frame chunks x 3 threads 10.85 k, 16 x 4 10.46 k (C2, same box)
    ap.add_argument("--e2e-chunk", type=int, default=0)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
