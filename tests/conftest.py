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


def parity_record(name, got, want, tol):
    """Append max |diff| / differing pixels / PSNR of a GPU-vs-reference comparison to gpurun_out/parity_r02.json
    (copied to profiles/ after the run: the north star asks for the tolerance stages' PSNR to be reported)."""
    import json
    d = np.abs(got.astype(np.int64) - want.astype(np.int64))
    mse = float(np.mean(d.astype(np.float64) ** 2))
    rec = {"test": name, "shape": list(got.shape), "tolerance_dn": tol, "max_abs_diff_dn": int(d.max()),
           "differing_px": int(np.count_nonzero(d)), "px": int(d.size),
           "psnr_db": None if mse == 0 else round(float(10 * np.log10(65535.0 ** 2 / mse)), 2)}
    path = os.path.join(ROOT, "gpurun_out", "parity_r02.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        recs = []
        if os.path.exists(path):
            with open(path) as f:
                recs = json.load(f)
        recs = [r for r in recs if r["test"] != name] + [rec]
        with open(path, "w") as f:
            json.dump(recs, f, indent=1)
    except OSError:
        pass
    print(f"{name}: max |diff| = {rec['max_abs_diff_dn']} DN, differing px = {rec['differing_px']} / {rec['px']}, PSNR = {rec['psnr_db']} dB")
    return rec


def reference_frames(tmp_path, clip, nframes, opts, with_headers=False):
    """Frames of <tmp_path>/<clip> from the unmodified reference's process_frame in a fresh process (tests/refrun.py).
    opts: dict with the fields of struct mlvfs (mlvfs.h:37-46)."""
    import subprocess
    order = ["chroma_smooth", "fix_bad_pixels", "fix_stripes", "dual_iso", "hdr_interpolation_method", "hdr_no_fullres",
             "hdr_no_alias_map", "fix_pattern_noise", "deflicker"]
    out = os.path.join(str(tmp_path), "ref_frames.npy")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "refrun.py"), str(tmp_path), clip, str(nframes), out]
                          + [str(int(opts.get(k, 0))) for k in order])
    if with_headers:
        return np.load(out), np.load(out + ".headers.npy")
    return np.load(out)
