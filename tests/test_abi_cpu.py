"""The C-ABI library loads without a GPU and exports every entry point include/resolv_b200.h declares; the product
path refuses to run without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

from resolv_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "resolv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rsv_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    names = _header_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert lib.rsv_abi_version() == 1


def test_python_binding_lists_the_header():
    assert set(_abi.ABI_SYMBOLS) <= set(_header_functions())


def test_structs_match_the_header_layout():
    import ctypes as C
    assert C.sizeof(_abi.Config) == 15 * 4
    assert C.sizeof(_abi.Counts) == 7 * 8 + 8
    assert _abi.HIT_DTYPE.itemsize == 32 and _abi.CONTACT_DTYPE.itemsize == 32 and _abi.RAY_HIT_DTYPE.itemsize == 64
    assert C.sizeof(_abi.RayCounts) == 4 * 8 + 16


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_abi.RsvError) as e:
        _abi.DeviceSpace(640, 360, 16, 16, max_shapes=16, max_verts=64)
    assert e.value.code == _abi.RSV_ERR_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "resolv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "liboracle" not in txt and "resolv_oracle" not in txt, f"{f} links the oracle"
