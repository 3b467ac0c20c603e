import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import mlvfs_b200
        return mlvfs_b200.lib().mlvb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a `-m gpu` run on a machine without a device must fail loudly, not skip: there is no CPU path
    pass


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.load_oracle()
    return pyoracle


@pytest.fixture(scope="session")
def ref(oracle):
    lib = oracle.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref (compiled reference) not built on this machine")
    return lib


@pytest.fixture(scope="session")
def gpu_ctx():
    import mlvfs_b200
    ctx = mlvfs_b200.Context(device=0, slots=4)
    yield ctx
    ctx.close()


@pytest.fixture()
def fresh_ctx(gpu_ctx):
    """Session context with per-clip state cleared and the dither stream re-seeded (fresh process = seed 1)."""
    gpu_ctx.reset_clip_state()
    gpu_ctx.seed_dither(1)
    import mlvfs_b200
    mlvfs_b200.lib().free_focus_pixel_maps()
    mlvfs_b200.lib().stripes_free_corrections()
    # the drop-in symbols run on the process-wide default context: clear its per-clip / dual-ISO state too
    mlvfs_b200.Context.default().reset_clip_state()
    return gpu_ctx
