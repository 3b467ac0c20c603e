"""The host LZMA1 decoder for legacy MLV_VIDEO_CLASS_FLAG_LZMA payloads (mlvfs_b200/csrc/lzma_dec.cu, plain host
C++, built here with g++): against Python's lzma streams, and -- where the compiled reference exists -- the
reference's get_image_data (main.c:598-616, LZMA SDK) must accept the very same synthetic payloads."""
import ctypes as C
import lzma
import os
import subprocess

import numpy as np
import pytest

from mlvfs_b200 import mlvformat as F, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def decode(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lzma") / "liblzma_dec.so")
    subprocess.check_call(["g++", "-O2", "-x", "c++", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "mlvfs_b200", "csrc", "lzma_dec.cu")])
    fn = C.CDLL(so)._Z16mlvb_lzma_decodePhPmPKhmS2_
    fn.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t, C.c_char_p]

    def run(stream, props, cap):
        out = (C.c_uint8 * max(cap, 1))()
        n = C.c_size_t(cap)
        rc = fn(out, C.byref(n), stream, len(stream), props)
        return rc, bytes(out)[:n.value]

    return run


@pytest.mark.parametrize("lc,lp,pb,dict_size", [(3, 0, 2, 1 << 16), (0, 2, 0, 4096), (4, 0, 4, 1 << 20), (2, 2, 3, 1 << 23)])
def test_lzma_decoder_matches_python_lzma(decode, lc, lp, pb, dict_size):
    rng = np.random.default_rng(lc * 7 + lp)
    data = (rng.integers(0, 40, 200000) + np.arange(200000) // 1000).astype(np.uint8).tobytes() + b"abc" * 5000 + bytes(3000)
    alone = lzma.compress(data, format=lzma.FORMAT_ALONE,
                          filters=[{"id": lzma.FILTER_LZMA1, "lc": lc, "lp": lp, "pb": pb, "dict_size": dict_size}])
    props, stream = alone[:5], alone[13:]
    for cap in (len(data), len(data) - 777, len(data) + 100, 1):
        rc, got = decode(stream, props, cap)
        assert rc == 0 and got == data[:cap]                    # a full buffer stops the decode (LZMA_FINISH_ANY); the end marker too
    rc, _ = decode(stream[:len(stream) // 2], props, len(data))
    assert rc < 0                                               # truncated stream
    rc, _ = decode(b"\x01" + stream[1:], props, len(data))
    assert rc < 0                                               # a range-coded stream starts with a zero byte
    rc, _ = decode(stream, bytes([225]) + props[1:], len(data))
    assert rc < 0                                               # lc/lp/pb byte out of range


def test_frame_payload_round_trip(decode):
    img = synth.make_frame(320, 180, 3)
    pl = synth.lzma_payload(img).tobytes()
    size = int(np.frombuffer(pl[:4], "<u4")[0])
    rc, got = decode(pl[9:], pl[4:9], size)
    assert rc == 0 and got == synth.pack_bits(img).tobytes()


def test_reference_accepts_the_synthetic_lzma_clip(ref, oracle, tmp_path):
    """Pins the test input: the reference's own LZMA branch decodes our synthetic LZMA clip to the original frames."""
    w, h, n = 320, 180, 2
    hdr = F.make_frame_headers(w, h, video_class=F.VIDEO_CLASS_RAW | F.VIDEO_CLASS_FLAG_LZMA)
    frames = [synth.make_frame(w, h, i) for i in range(n)]
    synth.write_mlv(str(tmp_path / "Z.MLV"), (synth.lzma_payload(f).tobytes() for f in frames), hdr)
    ref.ref_set_mlv_dir(str(tmp_path).encode())
    ref.ref_set_options(0, 0, 0, 0, 0, 0, 0, 0, 0)
    for i in range(n):
        out = np.zeros((h, w), np.uint16)
        with oracle.quiet_stdout():
            assert ref.ref_process_frame(b"/Z.MLV/Z_%06d.dng" % i, out.ctypes.data_as(C.c_void_p), out.nbytes, None) == out.nbytes
        assert np.array_equal(out, frames[i])
