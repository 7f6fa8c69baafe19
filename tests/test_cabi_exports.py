"""CPU: libfmgpu.so builds for sm_100a, loads, and exports every symbol include/fm_gpu.h declares
(no compute calls without a GPU; compute entry points must fail loudly instead of falling back)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from find_motion_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_all_exported(lib):
    from find_motion_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "fm_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fm_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libfmgpu.so"


def test_struct_layouts_match_header():
    from find_motion_b200 import _lib
    assert C.sizeof(_lib.fm_frame_stats) == 32
    assert C.sizeof(_lib.fm_component) == 20
    assert C.sizeof(_lib.fm_config) == 72
    assert C.sizeof(_lib.fm_info) == 48


def test_flag_values_match_header():
    """The Python binding's FLAG_* constants are the header's FM_FLAG_* values (a maintainer's own stub copies the header)."""
    from find_motion_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "fm_gpu.h")).read()
    flags = dict(re.findall(r"#define\s+FM_FLAG_([A-Z_]+)\s+(\d+)", hdr))
    assert set(flags) >= {"KEEP_PLANES", "NO_FUSED", "NO_UMMA", "UMMA_APRON", "UMMA", "NO_ROWS"}
    for name, val in flags.items():
        assert getattr(_lib, "FLAG_" + name) == int(val), name
    vals = [int(v) for v in flags.values()]
    assert len(set(vals)) == len(vals) and all(v & (v - 1) == 0 for v in vals), "flags are distinct single bits"


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from find_motion_b200 import _lib
    from find_motion_b200.engine import MotionEngine
    with pytest.raises(_lib.FmError):
        MotionEngine(64, 48, box_size=64)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "find_motion_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn
