#!/usr/bin/env python3
"""bench.py -- DNG frames/s of the MLVFS per-frame raw path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]           our CUDA path
  python bench.py --impl reference [...]                        the reference's CPU path (oracle/_ref)
  torchrun --nproc-per-node N bench.py --gpus N ...             one rank per GPU, frames sharded by rank

Workload (config.workload): BASELINE.json configs[1] -- 1920x1080 14-bit uncompressed MLV frames with
--stripes --bad-pix --cs3x3 (the full single-ISO correction chain), synthetic input (mlvfs_b200/synth.py).
A "step" is one pass of the hot path over a batch of `frames_per_step` frames.

  value  frames/s with the packed payloads already resident in HBM (mlvb_process_batch_device),
         CUDA-event timed on the launching stream, max over ranks.
  e2e    frames/s through the host-buffer C ABI (mlvb_submit / mlvb_wait): pinned host payload ->
         H2D -> kernels -> D2H of the finished 16-bit frame, all inside the timed region.
  roofline      the dominant kernel (chroma smoothing + fused stripes store), algorithmic bytes per
                launch / its CUDA-event duration measured in the timed region, vs MEASURED_PEAKS.json.
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

W, H, BPP = 1920, 1080, 14
NPIX = W * H
OPTS = dict(chroma_smooth=3, fix_bad_pixels=1, fix_stripes=1)
WORKLOAD = "C2: 1920x1080 14-bit uncompressed MLV, --stripes --bad-pix --cs3x3"
METRIC = "DNG frames/sec (1920x1080 14-bit, full single-ISO correction chain)"
ALGO_BYTES_PER_FRAME = NPIX * 14 // 8 + NPIX * 2          # C2 = 7 776 000 B
CHROMA_BYTES_PER_PX = 4                                   # chroma-smooth (16-bit) 2 B in + 2 B out


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU during the timed region (NVML)."""

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

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown"}

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


def make_inputs(frames_per_step, distinct=8):
    from mlvfs_b200 import synth
    base = [synth.pack_bits(synth.make_frame(W, H, i, hot_cold=True, stripes=True)) for i in range(distinct)]
    return np.stack([base[i % distinct] for i in range(frames_per_step)])          # [B, words] uint16


# ----------------------------------------------------------------------------------------------
# reference / CPU arm

def reference_runner():
    """Returns (kind, cores, fn(nframes) -> seconds) timing the reference CPU path on C2 frames."""
    from mlvfs_b200 import mlvformat as F, synth
    from oracle import pyoracle as O
    ref = O.load_ref()
    cores = os.cpu_count() or 1
    hdr = F.make_frame_headers(W, H)
    ri = hdr.rawi_hdr.raw_info
    nclip = 8
    frames = [synth.make_frame(W, H, i, hot_cold=True, stripes=True) for i in range(nclip)]
    if ref is not None:
        threads = min(cores, 32)
        tmp = tempfile.mkdtemp(prefix="mlvb_ref_")
        synth.write_mlv(os.path.join(tmp, "C2.MLV"), (synth.pack_bits(f).tobytes() for f in frames), hdr)
        ref.ref_set_mlv_dir(tmp.encode())
        ref.ref_set_options(OPTS["chroma_smooth"], OPTS["fix_bad_pixels"], OPTS["fix_stripes"], 0, 0, 0, 0, 0, 0)
        bufs = [np.empty(NPIX, np.uint16) for _ in range(threads)]

        def one(tid, idx):
            ref.ref_process_frame(b"/C2.MLV/C2_%06d.dng" % (idx % nclip), bufs[tid].ctypes.data_as(C.c_void_p),
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

    # oracle port (single thread per frame, frames spread over a few threads)
    threads = min(cores, 16)
    state = {}
    O.single_iso_chain(frames[:1], ri.black_level, ri.white_level, ri.frame_size, chroma_smooth_method=3,
                       fix_bad_pixels=1, fix_stripes=1, state=state)

    def run(nframes):
        counter = iter(range(nframes))
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    i = next(counter, None)
                if i is None:
                    return
                O.single_iso_chain([frames[i % nclip]], ri.black_level, ri.white_level, ri.frame_size,
                                   chroma_smooth_method=3, fix_bad_pixels=1, fix_stripes=1, state=state)

        ts = [threading.Thread(target=worker) for _ in range(threads)]
        t0 = time.perf_counter()
        [t.start() for t in ts]
        [t.join() for t in ts]
        return time.perf_counter() - t0

    return "port", threads, run


def cpu_baseline(budget_s=12.0):
    kind, cores, run = reference_runner()
    n = max(cores, 4)
    dt = run(n)                                   # calibration pass doubles as warm-up
    total = int(min(max(n, n * budget_s / max(dt, 1e-3)), 50 * cores))
    dt = run(total)
    return {"value": total / dt, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": f"{total} frames of the C2 workload through "
                      f"{'oracle/_ref process_frame (unmodified reference, gcc -O2)' if kind == 'reference' else 'the oracle port'}"
                      f" on {cores} threads, {dt:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, cores, run = reference_runner()
    per_step = max(cores * 2, 8)
    for _ in range(max(1, min(args.warmup, 1))):
        run(per_step)
    t = 0.0
    for _ in range(args.steps):
        t += run(per_step)
    fps = per_step * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": per_step},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{per_step} frames/step x {args.steps} steps on {cores} host threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm

def run_ours(args):
    import torch
    import mlvfs_b200 as M
    from mlvfs_b200 import mlvformat as F

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

    B = args.frames_per_step
    hdr = F.make_frame_headers(W, H)
    opts = M.Options(**OPTS)
    packed = make_inputs(B)
    stride = packed.shape[1] * 2
    ctx = M.Context(device=local, slots=args.slots)
    d_in = torch.from_numpy(packed.view(np.int16)).cuda()
    d_out = torch.empty((B, NPIX), dtype=torch.int16, device="cuda")
    stream = torch.cuda.Stream()

    def step():
        ctx.process_batch_device(hdr, opts, "bench_C2.MLV", d_in.data_ptr(), stride, stride, d_out.data_ptr(), NPIX, B,
                                 stream.cuda_stream)

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
    pin_in.array[:] = packed.view(np.uint8).reshape(-1)
    depth = args.slots
    pin_out = [M.PinnedBuffer(NPIX * 2) for _ in range(depth)]

    def e2e_step():
        q = collections.deque()
        for f in range(B):
            if len(q) == depth:
                ctx.wait(q.popleft())
            tk = ctx.submit(hdr, C.c_void_p(pin_in.ptr + f * stride), stride, opts, "bench_C2.MLV",
                            C.c_void_p(pin_out[f % depth].ptr))
            if tk < 0:
                raise RuntimeError(f"mlvb_submit failed: {tk}")
            q.append(tk)
        while q:
            ctx.wait(q.popleft())

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda")
    if dist is not None:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = B * args.steps * world / float(t.item())
    checksum = int(pin_out[0].array.view(np.uint16)[::4099].astype(np.uint64).sum())

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        roof = None
        if "chroma" in stages:
            tot_ms, spans = stages["chroma"]
            per_launch_s = tot_ms * 1e-3 / spans
            achieved = CHROMA_BYTES_PER_PX * NPIX * B / per_launch_s / 1e9
            roof = {"bound": "hbm", "kernel": "chroma_smooth_kernel<u16,3x3> + fused stripes store",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": peak_src, "launch_ms": per_launch_s * 1e3,
                    "algorithmic_bytes_per_launch": CHROMA_BYTES_PER_PX * NPIX * B,
                    "stage_ms_per_step": {k: v[0] / args.steps for k, v in stages.items()},
                    "chain_achieved_gbs": ALGO_BYTES_PER_FRAME * value / world / 1e9,
                    "chain_frac": ALGO_BYTES_PER_FRAME * value / world / 1e9 / peak}
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": W_,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16 pixels / int32 EV-LUT arithmetic", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": B, "sharding": f"frames by rank, {world} rank(s), no collective",
                       "cache": f"inputs+outputs per step {B * ALGO_BYTES_PER_FRAME / 1e6:.0f} MB > 126 MB L2 (no flush needed)"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": B * stride, "d2h_bytes_per_step": B * NPIX * 2,
                    "frames_in_flight": depth, "checksum": checksum},
            "gpu_launches": launches,
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
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
    ap.add_argument("--frames-per-step", type=int, default=256)
    ap.add_argument("--slots", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
