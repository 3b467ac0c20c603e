"""N>1 host logic on CPU: frame sharding by rank and the max-over-ranks reduction that bench.py uses,
exercised with world_size 2 over gloo (no GPU, no data-path collective: frames are independent)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from mlvfs_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = sharding.frames_for_rank(37, rank, world)
# every frame is owned by exactly one rank
owned = torch.zeros(37, dtype=torch.int32)
owned[mine] = 1
dist.all_reduce(owned)
assert bool((owned == 1).all()), owned
# whole-job time = max over ranks; throughput = total frames / that
t = torch.tensor([0.5 + rank])
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert float(t) == 0.5 + world - 1
n = torch.tensor([len(mine)])
dist.all_reduce(n)
assert int(n) == 37
dist.barrier()
if rank == 0:
    print("OK", sharding.describe(world))
dist.destroy_process_group()
'''


def test_frame_sharding_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.check_output(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29531", str(script)], env=env, text=True, stderr=subprocess.STDOUT, timeout=240)
    assert "OK" in out, out


def test_in_process_chunk_dealing_matches_the_host_layer():
    """sharding.gpu_for_frame mirrors host/frame_builder.c (context_for_frame: (frame / chunk) % contexts): every frame on
    exactly one GPU, a look-ahead chunk never split, and every shard primed with the clip's frame 0."""
    import re
    from mlvfs_b200 import sharding
    src = open(os.path.join(ROOT, "mlvfs_b200", "host", "frame_builder.c")).read()
    assert re.search(r"g_ctx\[\(frame / g_chunk\) % g_nctx\]", src), "the C dealing changed: update mlvfs_b200/sharding.py"
    for nframes, chunk, g in [(37, 4, 8), (256, 8, 8), (5, 16, 2), (64, 1, 3)]:
        seen = [0] * nframes
        for gpu in range(g):
            for a, b in sharding.chunks_for_gpu(nframes, gpu, chunk, g):
                assert a % chunk == 0 and b - a <= chunk
                for n in range(a, b):
                    assert sharding.gpu_for_frame(n, chunk, g) == gpu
                    seen[n] += 1
        assert seen == [1] * nframes
    assert sharding.prime_frames(3, 8) == [0]
