"""Frame sharding across GPUs: frames are independent, so rank r of G owns frames
r, r+G, r+2G, ... of a clip and no data-path collective exists.  Per-clip state is derived from frame
0 on every rank independently (a few KB of statistics; identical on all ranks because the input
frame and the dither stream are identical)."""


def frames_for_rank(nframes, rank, world):
    return list(range(rank, nframes, world))


def describe(world):
    return f"frame n -> rank n mod {world}; no collective; per-clip state recomputed from frame 0 on each rank"
