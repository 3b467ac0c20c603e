"""CPU-side checks of the drop-in boundary: struct layouts and exported symbols (no compute calls)."""
import ctypes as C
import os
import re

import pytest

import mlvfs_b200 as M
from mlvfs_b200 import mlvformat as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header_text):
    text = re.sub(r"/\*.*?\*/", "", header_text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", text)
    return sorted(set(n for n in names if n not in ("defined", "sizeof")))


def test_library_loads_and_exports_every_declared_symbol():
    lib = M.lib()
    with open(os.path.join(ROOT, "include", "mlvfs_b200.h")) as f:
        declared = _declared_functions(f.read())
    assert "mlvb_process_frame" in declared and "cr2hdr20_convert_data" in declared
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, f"declared in include/mlvfs_b200.h but not exported: {missing}"
    assert sorted(M.ABI_SYMBOLS) == declared, "mlvfs_b200.ABI_SYMBOLS out of sync with the header"


def test_only_the_abi_is_exported():
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", M.LIB_PATH], text=True)
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    extra = [s for s in syms if s not in M.ABI_SYMBOLS]
    assert not extra, extra


def test_frame_headers_layout_matches_header_file():
    """include/mlvb_mlv_format.h compiled by gcc must agree with the ctypes mirror."""
    import subprocess
    import tempfile
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "mlvfs_b200.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(struct frame_headers), sizeof(struct raw_info),
               offsetof(struct frame_headers, vidf_hdr), offsetof(struct frame_headers, file_hdr),
               offsetof(struct frame_headers, idnt_hdr), offsetof(struct frame_headers, rawi_hdr),
               offsetof(struct frame_headers, wbal_hdr), sizeof(mlvb_options), sizeof(mlvb_frame_result));
        return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "t.c"), "w") as f:
            f.write(src)
        subprocess.check_call(["gcc", "-std=gnu99", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"),
                               os.path.join(d, "t.c")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")], text=True).split()]
    FH = F.FrameHeaders
    want = [C.sizeof(FH), C.sizeof(F.RawInfo), FH.vidf_hdr.offset, FH.file_hdr.offset, FH.idnt_hdr.offset,
            FH.rawi_hdr.offset, FH.wbal_hdr.offset, C.sizeof(M.Options), C.sizeof(M.FrameResult)]
    assert got == want


def test_frame_headers_layout_matches_reference(ref):
    """Pin the layout against the compiled reference's own structs (mlvfs.h:51-63, raw.h:166-207)."""
    FH = F.FrameHeaders
    assert ref.ref_sizeof_frame_headers() == C.sizeof(FH)
    names = ["fileNumber", "position", "vidf_hdr", "file_hdr", "rtci_hdr", "idnt_hdr", "rawi_hdr", "expo_hdr",
             "lens_hdr", "wbal_hdr"]
    for i, n in enumerate(names):
        assert ref.ref_offsetof_frame_headers(i) == getattr(FH, n).offset, n
    assert ref.ref_offsetof_frame_headers(10) == FH.rawi_hdr.offset + F.RawiHdr.raw_info.offset
    RI = F.RawInfo
    for code, field in [(11, "black_level"), (12, "white_level"), (13, "bits_per_pixel"), (14, "frame_size"),
                        (16, "exposure_bias"), (17, "active_area")]:
        assert ref.ref_offsetof_frame_headers(code) == getattr(RI, field).offset, field
    assert ref.ref_offsetof_frame_headers(15) == C.sizeof(RI)


def test_no_cpu_fallback_without_device():
    """On a machine without a GPU the product must fail loudly, not compute on the CPU."""
    lib = M.lib()
    if lib.mlvb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError):
        M.Context(device=0)


def test_product_never_imports_the_oracle():
    """Nothing under mlvfs_b200/ may reference oracle/ (the oracle is test infrastructure)."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mlvfs_b200")):
        for fn in files:
            if fn.endswith((".py", ".c", ".h", ".cu", ".cuh", "Makefile")):
                with open(os.path.join(dirpath, fn), errors="ignore") as f:
                    t = f.read()
                if re.search(r"(from|import)\s+oracle|oracle/|liboracle|libmlvfs_ref|pyoracle", t):
                    bad.append(os.path.join(dirpath, fn))
    assert not bad, bad
