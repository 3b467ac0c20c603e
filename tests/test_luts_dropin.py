"""The exported EV-table accessors get_raw2ev / get_raw2evf / get_ev2raw (reference mlvfs.h:90-92, main.c:128-196),
entry for entry against the oracle's tables (which tests/test_oracle_vs_ref.py pins to the compiled reference) and,
where it is present, against the reference itself.  The product builds them on the host with libm and uploads
these very arrays to the device, so no GPU is needed to compare them."""
import ctypes as C

import numpy as np
import pytest

import mlvfs_b200 as M

EV = 32768


def _tables(lib, names, black):
    n = 16384 + black
    raw2ev = np.ctypeslib.as_array(getattr(lib, names[0])(black), (n,))
    raw2evf = np.ctypeslib.as_array(getattr(lib, names[1])(black), (n,))
    base = C.cast(C.addressof(getattr(lib, names[2])().contents) - 4 * 10 * EV, C.POINTER(C.c_int))
    return raw2ev, raw2evf, np.ctypeslib.as_array(base, (24 * EV,))


@pytest.mark.parametrize("black", [0, 1, 1024, 2048, 2049, 16384])
def test_dropin_luts_equal_the_oracle_entry_for_entry(oracle, black):
    orc = oracle.load_oracle()
    got = _tables(M.lib(), ("get_raw2ev", "get_raw2evf", "get_ev2raw"), black)
    want = _tables(orc, ("orc_raw2ev", "orc_raw2evf", "orc_ev2raw"), black)
    assert np.array_equal(got[0], want[0])
    assert got[0][black] == -2**31 and (black == 0 or got[0][black - 1] == 0) and got[0][black + 1] == 0 and got[0][black + 2] == EV
    assert np.array_equal(got[1][black + 1:], want[1][black + 1:])            # fp64 table: below / at black is 0 / -inf
    assert np.array_equal(got[2], want[2])
    assert got[2][10 * EV] == 1 and got[2][24 * EV - 1] == 16383 and got[2][10 * EV - 1] == 0


def test_dropin_luts_equal_the_reference(ref):
    for black in (0, 2048, 16384):
        got = _tables(M.lib(), ("get_raw2ev", "get_raw2evf", "get_ev2raw"), black)
        want = _tables(ref, ("get_raw2ev", "get_raw2evf", "get_ev2raw"), black)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2])
        assert np.array_equal(got[1][black + 1:], want[1][black + 1:])


def test_black_level_above_the_table_is_refused():
    L = M.lib()
    assert not L.get_raw2ev(16385) and not L.get_raw2evf(20000)               # main.c:131-135, 157-161: NULL + message
