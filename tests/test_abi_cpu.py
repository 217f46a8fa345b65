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


def test_flag_constants_match_the_header():
    src = open(os.path.join(ROOT, "include", "resolv_b200.h")).read()
    defs = {m.group(1): int(m.group(2), 16) for m in re.finditer(r"#define\s+(RSV_[A-Z_0-9]+)\s+0x([0-9a-fA-F]+)u", src)}
    for name, val in (("RSV_META_STATIC", _abi.META_STATIC), ("RSV_META_ABSENT", _abi.META_ABSENT), ("RSV_META_TESTER", _abi.META_TESTER),
                      ("RSV_FRAME_INCREMENTAL", _abi.FRAME_INCREMENTAL), ("RSV_FRAME_GRID_ONLY", _abi.FRAME_GRID_ONLY),
                      ("RSV_FRAME_NO_TAG_FILTERS", _abi.FRAME_NO_TAG_FILTERS), ("RSV_FRAME_TESTER_TABLE", _abi.FRAME_TESTER_TABLE),
                      ("RSV_FRAME_CANDIDATE_WALK", _abi.FRAME_CANDIDATE_WALK),
                      ("RSV_LINE_DDA", _abi.LINE_DDA), ("RSV_LINE_HOME_ONLY", _abi.LINE_HOME_ONLY)):
        assert defs[name] == val, name
    frame_bits = [v for k, v in defs.items() if k.startswith(("RSV_FRAME_", "RSV_LINE_"))]
    assert len(frame_bits) == len(set(frame_bits)), "frame / line flags share one word and must not collide"


def test_reference_arm_of_the_platformer_batch_runs_on_the_host():
    """bench.py --impl reference --workload cfg4 (the oracle port on worker processes), at a size that takes seconds."""
    import json
    import subprocess
    import sys
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg4",
                                   "--spaces", "8", "--steps", "1", "--warmup", "0", "--cpu-threads", "2"], text=True, timeout=300)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["cores"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "configs[3]" in line["config"]["workload"]
