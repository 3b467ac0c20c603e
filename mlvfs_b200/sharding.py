"""Frame sharding across GPUs: frames are independent, so no data-path collective exists.

Two dealings, both used by the product:
  * across processes (torchrun, one rank per GPU; bench.py): rank r of G owns frames r, r+G, r+2G, ... of the clip;
  * inside one process (host/frame_builder.c: frame_builder_use_gpus): frames are dealt in chunks of `chunk`
    consecutive frames, frame n -> GPU (n // chunk) % G, so that a look-ahead chunk is one device batch on one GPU.
Per-clip state (bad-pixel map, stripe coefficients, EV tables) is derived from the clip's frame 0 on every rank / GPU
independently -- identical everywhere because the input frame and the dither stream are identical -- so every shard is
primed with frame 0 before its own frames (bench.py does that through prime_frames())."""


def frames_for_rank(nframes, rank, world):
    """Frame indices of the clip that rank `rank` of `world` processes."""
    return list(range(rank, nframes, world))


def gpu_for_frame(n, chunk, ngpus):
    """The in-process dealing of host/frame_builder.c (context_for_frame): chunked round-robin."""
    return (n // max(chunk, 1)) % max(ngpus, 1)


def chunks_for_gpu(nframes, gpu, chunk, ngpus):
    """[(first, last + 1)] frame ranges GPU `gpu` builds: the look-ahead chunks that land on it."""
    chunk = max(chunk, 1)
    return [(c, min(c + chunk, nframes)) for c in range(0, nframes, chunk) if gpu_for_frame(c, chunk, ngpus) == gpu]


def prime_frames(rank, world):
    """Frames a shard must push through before its own so that its per-clip state equals every other shard's: the clip's
    frame 0 (the reference creates the state from the first frame it processes)."""
    return [0]


def describe(world):
    return f"frame n -> rank n mod {world}; no collective; per-clip state recomputed from frame 0 on each rank"
