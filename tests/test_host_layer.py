"""Host-side C layer (mlvfs_b200/host): MLV index cache and frame cache + prefetch queue.  CPU only."""
import ctypes as C
import os
import subprocess
import threading
import time

import numpy as np
import pytest

from mlvfs_b200 import mlvformat as F, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "mlvfs_b200", "libmlvfs_b200_host.so")


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(HOST_SO):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mlvfs_b200", "host")])
    lib = C.CDLL(HOST_SO)
    lib.mlv_get_frame_headers.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
    lib.mlv_get_frame_count.argtypes = [C.c_char_p]
    lib.get_or_create_image_buffer.restype = C.c_void_p
    lib.get_or_create_image_buffer.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.release_image_buffer_by_path.argtypes = [C.c_char_p]
    lib.resource_manager_set_prefetch.argtypes = [C.c_int, C.c_int, C.c_void_p]
    lib.resource_manager_prefetch_stats.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    return lib


def _write_clip_with_metadata(path, w, h, nframes):
    """MLV whose EXPO/LENS/WBAL blocks change mid-clip and whose blocks are stored out of timestamp order."""
    hdr = F.make_frame_headers(w, h)
    frames = [synth.make_frame(w, h, i) for i in range(nframes)]
    blocks = []   # (timestamp, bytes)

    def expo(ts, iso):
        e = F.ExpoHdr()
        F._tag(e.blockType, "EXPO")
        e.blockSize = C.sizeof(F.ExpoHdr)
        e.timestamp = ts
        e.isoValue = iso
        e.shutterValue = 20000 + iso
        return bytes(e)

    def lens(ts, fl):
        l = F.LensHdr()
        F._tag(l.blockType, "LENS")
        l.blockSize = C.sizeof(F.LensHdr)
        l.timestamp = ts
        l.focalLength = fl
        F._tag(l.lensName, "EF50mm")
        return bytes(l)

    def wbal(ts, k):
        b = F.WbalHdr()
        F._tag(b.blockType, "WBAL")
        b.blockSize = C.sizeof(F.WbalHdr)
        b.timestamp = ts
        b.kelvin = k
        return bytes(b)

    def null(ts):
        n = F.MlvHdr()
        F._tag(n.blockType, "NULL")
        n.blockSize = 64
        n.timestamp = ts
        return bytes(n) + b"\0" * (64 - C.sizeof(F.MlvHdr))

    rawi = F.RawiHdr.from_buffer_copy(bytes(hdr.rawi_hdr))
    rawi.timestamp = 1
    idnt = F.IdntHdr.from_buffer_copy(bytes(hdr.idnt_hdr))
    idnt.timestamp = 2
    blocks.append((1, bytes(rawi)))
    blocks.append((2, bytes(idnt)))
    blocks.append((3, expo(3, 100)))
    blocks.append((4, lens(4, 50)))
    blocks.append((5, wbal(5, 5600)))
    for i, fr in enumerate(frames):
        ts = 1000 + i * 1000
        v = F.VidfHdr.from_buffer_copy(bytes(hdr.vidf_hdr))
        v.frameNumber = i
        v.timestamp = ts
        v.frameSpace = 8 * (i % 3)
        payload = synth.pack_bits(fr).tobytes()
        v.blockSize = C.sizeof(F.VidfHdr) + v.frameSpace + len(payload)
        blocks.append((ts, bytes(v) + b"\0" * v.frameSpace + payload))
        if i % 3 == 1:
            blocks.append((ts + 10, expo(ts + 10, 200 + i)))
        if i % 4 == 2:
            blocks.append((ts + 20, lens(ts + 20, 24 + i)))
            blocks.append((ts + 30, null(ts + 30)))
    # store out of order: swap some neighbours so that the timestamp sort matters
    order = list(range(len(blocks)))
    for k in range(6, len(order) - 1, 5):
        order[k], order[k + 1] = order[k + 1], order[k]
    fh = F.FileHdr.from_buffer_copy(bytes(hdr.file_hdr))
    fh.videoFrameCount = nframes
    with open(path, "wb") as f:
        f.write(bytes(fh))
        for k in order:
            f.write(blocks[k][1])
    return hdr, frames


def test_index_cache_matches_reference_header_walk(host, ref, tmp_path):
    clip = str(tmp_path / "META.MLV")
    hdr, frames = _write_clip_with_metadata(clip, 64, 32, 14)
    assert host.mlv_get_frame_count(clip.encode()) == ref.ref_get_frame_count(clip.encode()) == 14
    for i in range(14):
        ours, theirs = F.FrameHeaders(), F.FrameHeaders()
        assert host.mlv_get_frame_headers(clip.encode(), i, C.byref(ours)) == 1
        assert ref.ref_get_frame_headers(clip.encode(), i, C.byref(theirs)) == 1
        assert bytes(ours) == bytes(theirs), f"frame {i}: headers differ"
    bad = F.FrameHeaders()
    assert host.mlv_get_frame_headers(clip.encode(), 14, C.byref(bad)) == 0


def test_index_is_rebuilt_when_the_clip_changes_on_disk(host, tmp_path):
    """The reference walks the file on every request (main.c:429-558), so a clip that is re-recorded or still
    growing is always seen as it is; the cached index has to notice the change (size / mtime of the first chunk)."""
    clip = str(tmp_path / "GROW.MLV")
    _write_clip_with_metadata(clip, 64, 32, 6)
    assert host.mlv_get_frame_count(clip.encode()) == 6
    time.sleep(0.02)
    hdr, _ = _write_clip_with_metadata(clip, 64, 32, 11)
    assert host.mlv_get_frame_count(clip.encode()) == 11
    got = F.FrameHeaders()
    assert host.mlv_get_frame_headers(clip.encode(), 10, C.byref(got)) == 1
    assert host.mlv_get_frame_headers(clip.encode(), 11, C.byref(got)) == 0
    assert host.mlv_get_frame_count(str(tmp_path / "MISSING.MLV").encode()) == 0


class ImageBuffer(C.Structure):
    pass


ImageBuffer._fields_ = [("next", C.POINTER(ImageBuffer)), ("dng_filename", C.c_char_p), ("header_size", C.c_size_t),
                        ("size", C.c_size_t), ("header", C.c_void_p), ("data", C.c_void_p),
                        ("mutex", C.c_byte * 40), ("in_use", C.c_int), ("prefetched", C.c_int), ("pins", C.c_int)]
CBR = C.CFUNCTYPE(C.c_int, C.POINTER(ImageBuffer))


def test_frame_cache_and_prefetch_queue(host):
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    built = []
    lock = threading.Lock()

    def build(ib):
        name = ib.contents.dng_filename.decode()
        time.sleep(0.01)
        with lock:
            built.append(name)
        ib.contents.size = 64
        ib.contents.data = libc.malloc(64)
        ib.contents.header = libc.malloc(16)
        ib.contents.header_size = 16
        return 1

    cbr = CBR(build)
    limit = C.CFUNCTYPE(C.c_int, C.c_char_p)(lambda p: 20)
    host.free_all_image_buffers()
    host.resource_manager_set_prefetch(4, 4, limit)
    created = C.c_int()
    try:
        for i in range(20):
            path = b"/clip.MLV/clip_%06d.dng" % i
            ib = host.get_or_create_image_buffer(path, cbr, C.byref(created))
            assert ib
            assert C.cast(ib, C.POINTER(ImageBuffer)).contents.data
            host.release_image_buffer_by_path(path)
            time.sleep(0.005)
        time.sleep(0.1)
        # every frame built exactly once, none beyond the clip's frame limit
        assert sorted(built) == sorted(set(built))
        assert set(built) == {f"/clip.MLV/clip_{i:06d}.dng" for i in range(20)}
        b, h = C.c_uint64(), C.c_uint64()
        host.resource_manager_prefetch_stats(C.byref(b), C.byref(h))
        assert b.value >= 10 and h.value >= 10, (b.value, h.value)
        assert host.get_image_buffer_count() <= 4 + 2 * (4 + 1) + 4
        # a cached frame is returned without calling the callback again
        n = len(built)
        path = b"/clip.MLV/clip_%06d.dng" % 19
        host.get_or_create_image_buffer(path, cbr, C.byref(created))
        assert created.value == 0 and len(built) == n
    finally:
        host.resource_manager_set_prefetch(0, 0, None)
        host.free_all_image_buffers()


BATCH_CBR = C.CFUNCTYPE(C.c_int, C.POINTER(C.POINTER(ImageBuffer)), C.c_int)


def test_prefetch_in_chunks_through_the_batch_builder(host):
    """Look-ahead frames are requested in chunks of `batch` consecutive frames; each chunk reaches the batch builder
    as ONE call holding only the frames that are not cached yet, every frame is built exactly once, and a reader that
    asks for a frame of a chunk under construction waits for it instead of building it again."""
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    host.resource_manager_set_batch_builder.argtypes = [C.c_void_p, C.c_int]
    host.resource_manager_prefetch_batches.restype = C.c_uint64
    singles, batches = [], []
    lock = threading.Lock()

    def fill(ib):
        ib.size = 64
        ib.data = libc.malloc(64)
        ib.header = libc.malloc(16)
        ib.header_size = 16

    def build(ib):
        with lock:
            singles.append(ib.contents.dng_filename.decode())
        time.sleep(0.004)
        fill(ib.contents)
        return 1

    def build_batch(bufs, n):
        names = [bufs[k].contents.dng_filename.decode() for k in range(n)]
        with lock:
            batches.append(names)
        time.sleep(0.02)
        for k in range(n):
            fill(bufs[k].contents)
        return 1

    cbr, bcbr = CBR(build), BATCH_CBR(build_batch)
    nframes = 50
    limit = C.CFUNCTYPE(C.c_int, C.c_char_p)(lambda p: nframes)
    host.free_all_image_buffers()
    before = host.resource_manager_prefetch_batches()
    host.resource_manager_set_batch_builder(bcbr, 8)
    host.resource_manager_set_prefetch(16, 0, limit)
    created = C.c_int()
    try:
        for i in range(nframes):
            path = b"/clip.MLV/clip_%06d.dng" % i
            ib = host.get_or_create_image_buffer(path, cbr, C.byref(created))
            assert ib and C.cast(ib, C.POINTER(ImageBuffer)).contents.data
            host.release_image_buffer_by_path(path)
        time.sleep(0.1)
        num = lambda s: int(s[-10:-4])
        flat = [f for b in batches for f in b]
        assert sorted(singles + flat) == [f"/clip.MLV/clip_{i:06d}.dng" for i in range(nframes)]     # each exactly once
        assert len(singles) <= 2                                   # frame 0 (nothing was ahead of it yet), rarely one more
        for b in batches:
            ns = [num(f) for f in b]
            assert ns == list(range(ns[0], ns[0] + len(ns)))       # consecutive frames
            assert ns[0] // 8 == ns[-1] // 8                       # of one chunk (= one GPU)
        assert max(len(b) for b in batches) == 8
        assert host.resource_manager_prefetch_batches() - before == len([b for b in batches if len(b) > 1])
    finally:
        host.resource_manager_set_prefetch(0, 0, None)
        host.resource_manager_set_batch_builder(None, 1)
        host.free_all_image_buffers()


def test_buffers_in_use_or_being_looked_up_are_never_evicted(host):
    """Readers hold frames (in_use) while a flood of other requests pushes the cache far beyond its soft limit: the
    held frames keep their data pointer; idle ones are evicted (resource_manager.c:195-227 frees by age only)."""
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]

    def build(ib):
        ib.contents.size = 64
        ib.contents.data = libc.malloc(64)
        C.memset(ib.contents.data, int(ib.contents.dng_filename[-5:-4]) + 1, 64)
        ib.contents.header = libc.malloc(16)
        ib.contents.header_size = 16
        return 1

    cbr = CBR(build)
    host.free_all_image_buffers()
    host.resource_manager_set_prefetch(0, 0, None)
    created = C.c_int()
    held = {}
    try:
        for i in range(6):                                         # more held frames than the soft limit of 4
            path = b"/hold.MLV/hold_%06d.dng" % i
            held[path] = C.cast(host.get_or_create_image_buffer(path, cbr, C.byref(created)), C.POINTER(ImageBuffer))
        stop = threading.Event()

        def flood(t):
            k = 0
            while not stop.is_set():
                path = b"/flood%d.MLV/flood_%06d.dng" % (t, k % 500)
                assert host.get_or_create_image_buffer(path, cbr, None)
                host.release_image_buffer_by_path(path)
                k += 1

        ts = [threading.Thread(target=flood, args=(t,)) for t in range(4)]
        [t.start() for t in ts]
        time.sleep(0.5)
        stop.set()
        [t.join() for t in ts]
        for path, ib in held.items():
            assert ib.contents.in_use == 1 and ib.contents.dng_filename == path
            data = C.string_at(ib.contents.data, 64)
            assert data == bytes([int(path[-5:-4]) + 1]) * 64
        assert host.get_image_buffer_count() <= 6 + 4 + 4          # held + soft limit + lookups in flight
        for path in held:
            host.release_image_buffer_by_path(path)
    finally:
        host.free_all_image_buffers()
