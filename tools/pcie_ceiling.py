#!/usr/bin/env python3
"""What the box's PCIe fabric can carry for this path: every rank copies frame-sized pinned buffers host->device and
device->host at the same time (two streams, buffers sized like BASELINE config 2: 3.63 MB packed in, 4.15 MB out),
no kernels.  Prints one JSON line: aggregate GB/s per direction and the frames/s that bandwidth would allow --
the ceiling the end-to-end numbers of bench.py are compared with (profiles/r02_pcie_ceiling.json).

    python tools/pcie_ceiling.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_ceiling.py
"""
import json
import os
import time

import torch


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {}
    for label, in_bytes, out_bytes, per_call in (("C2 frame by frame", 3628800, 4147200, 1), ("C2 32 frames per copy", 3628800 * 32, 4147200 * 32, 32),
                                                 ("C4 frame by frame", 32659200, 37324800, 1)):
        nbuf = 4
        h_in = [torch.empty(in_bytes, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
        h_out = [torch.empty(out_bytes, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
        d_in = [torch.empty(in_bytes, dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
        d_out = [torch.empty(out_bytes, dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

        def burst(n):
            for i in range(n):
                with torch.cuda.stream(s_up):
                    d_in[i % nbuf].copy_(h_in[i % nbuf], non_blocking=True)
                with torch.cuda.stream(s_dn):
                    h_out[i % nbuf].copy_(d_out[i % nbuf], non_blocking=True)

        burst(8)
        torch.cuda.synchronize()
        n = max(16, int(2e9 / (in_bytes + out_bytes)))          # about 2 GB per direction-pair per rank
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        burst(n)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        out[label] = {"h2d_gbs": world * n * in_bytes / dt / 1e9, "d2h_gbs": world * n * out_bytes / dt / 1e9,
                      "frames_per_s": world * n * per_call / dt, "copies_per_rank": n, "seconds": dt}
        del h_in, h_out, d_in, d_out
    if rank == 0:
        print(json.dumps({"what": "simultaneous pinned H2D + D2H, no kernels, max time over ranks", "n_gpus": world,
                          "host_cores": os.cpu_count(), "ceilings": out}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
