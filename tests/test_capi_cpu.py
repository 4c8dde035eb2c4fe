"""The C-ABI library loads and exports exactly the symbols include/ntm_b200.h declares; argument validation and
error strings work without a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
import ntm_b200
from ntm_b200 import lib

HEADER = os.path.join(ROOT, "include", "ntm_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ntm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 12
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in include/ntm_b200.h but not exported"
    assert sorted(lib.SIGNATURES) == names, "lib.py binds a different symbol set than the header declares"


def test_query_and_strerror():
    L = lib.load()
    assert L.ntm_query(lib.Q_VERSION) == 2
    assert L.ntm_query(lib.Q_MODE_MASK) & 1
    assert L.ntm_query(12345) == -1
    assert L.ntm_strerror(0) == b"ok"
    assert b"delay" in L.ntm_strerror(-5)
    assert L.ntm_set_tuning(-1, 0) == -1 and L.ntm_set_tuning(0, 0) == 0


def test_check_maps_errors():
    with pytest.raises(AssertionError):
        lib.check(-5)
    with pytest.raises(RuntimeError):
        lib.check(-1)
    lib.check(0)


def test_argument_validation_without_device():
    L = lib.load()
    h = ctypes.c_void_p()
    assert L.ntm_gru_prepare(None, None, None, None, None, None, 64, 0, ctypes.byref(h)) == -1
    buf = (ctypes.c_float * (192 * 64))()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.ntm_gru_prepare(p, p, p, p, p, None, 65, 0, ctypes.byref(h)) == -2     # NTM_EUNSUPPORTED: 1 <= H <= 64 is built
    assert L.ntm_gru_prepare(p, p, p, p, p, None, 0, 0, ctypes.byref(h)) == -2
    if not torch.cuda.is_available():
        assert L.ntm_gru_prepare(p, p, p, p, p, None, 64, 0, ctypes.byref(h)) == -6  # no device, no fallback
    assert L.ntm_gru_forward(None, 0, None, 0, None, 0, None, None, 1, 1, 0, None) == -1   # bad handle
    assert L.ntm_delay_forward(None, 0, None, 0, None, 0, None, None, -1, 1, 1, 0, 0, None) == -1
    L.ntm_destroy(None)                                                                 # no-op


def test_no_cpu_fallback():
    m = ntm_b200.RNN(1, 64, 1, False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 1, 16))
    d = ntm_b200.TimeVaryingDelayLine(max_delay=4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        d(torch.zeros(2, 1, 16), torch.zeros(2, 1, 16))
