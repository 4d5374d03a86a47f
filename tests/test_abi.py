"""CPU tests of the drop-in boundary: libmct_b200.so builds for sm_100a, loads, exports every symbol
include/mct_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mct_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mct_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from multicameratracking_b200 import capi
    lib = capi.load()
    declared = _header_symbols()
    assert sorted(capi.SYMBOLS) == declared
    for s in declared:
        assert hasattr(lib, s), s


def test_library_is_sm100a_only():
    import subprocess
    from multicameratracking_b200 import capi
    capi.load()
    out = subprocess.run(["cuobjdump", "--list-elf", capi.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from multicameratracking_b200 import capi
    with pytest.raises(capi.MctError) as e:
        capi.Context(0)
    assert e.value.code == capi.E_CUDA


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multicameratracking_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_py" not in txt and "mct_oracle" not in txt and "libmct_oracle" not in txt, f
