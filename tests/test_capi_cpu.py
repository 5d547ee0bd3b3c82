"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/hector_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import pytest

from hector_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "hector_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hx_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_capi.EXPORTS)


def test_library_exports_every_symbol():
    if not os.path.exists(_capi.lib_path()):
        pytest.skip("libhector_b200.so not built (run __graft_entry__.build())")
    L = C.CDLL(_capi.lib_path())
    for s in header_symbols():
        assert hasattr(L, s), s
    _capi.lib()
    assert b"sm_100a" in _capi.lib().hx_version()


def test_no_cpu_fallback():
    """without a CUDA device the engine must refuse to exist"""
    if not os.path.exists(_capi.lib_path()):
        pytest.skip("libhector_b200.so not built")
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _capi.lib()
    cfg = _capi.HxConfig(4, 1, 1745, 2300, 0, 0)
    h = C.c_void_p()
    rc = L.hx_create(C.byref(cfg), C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU path" in L.hx_last_error(None) or b"CUDA" in L.hx_last_error(None)


def test_bad_arguments_rejected_without_gpu():
    if not os.path.exists(_capi.lib_path()):
        pytest.skip("libhector_b200.so not built")
    L = _capi.lib()
    cfg = _capi.HxConfig(0, 1, 1745, 2300, 0, 0)
    h = C.c_void_p()
    assert L.hx_create(C.byref(cfg), C.byref(h)) == -1
    assert L.hx_destroy(None) == 0
