"""CPU: the C-ABI library is built, loads, and exports every symbol include/lav_b200.h declares."""
import ctypes
import os
import re

from lav_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "lav_b200.h")).read()
    declared = set(re.findall(r"\b(lavb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    handle = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(handle, s)]
    assert not missing, f"declared but not exported: {missing}"
    assert declared == set(capi.exported_symbols()), declared ^ set(capi.exported_symbols())


def test_abi_version_and_error_string():
    lib = capi.lib()
    assert lib.lavb_abi_version() == 3
    assert isinstance(lib.lavb_last_error(), bytes)


def test_no_cpu_fallback():
    import pytest
    import torch
    from lav_b200.lidar import LiDARModel
    m = LiDARModel(16, [64, 64], "cnn").eval()
    with pytest.raises(capi.LavbError):
        m([torch.zeros(10, 11)], [10])
