"""CPU tests of the C-ABI library: it loads, exports every symbol the header declares, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

import oracle_lib as ol

LIB = os.path.join(ol.ROOT, "taxator-tk_b200", "lib", "libtaxator_rpa_b200.so")
HEADER = os.path.join(ol.ROOT, "include", "taxator_rpa_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(trpa_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    assert os.path.exists(LIB), "build the extension first: make -C taxator-tk_b200"
    L = ctypes.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), s
    assert L.trpa_abi_version() == 9


def test_binding_lists_all_exports():
    import rpa_b200
    assert sorted(rpa_b200.EXPORTS) == declared_symbols()


def test_struct_layouts():
    import rpa_b200
    assert rpa_b200.CAND_DTYPE.itemsize == 36
    assert rpa_b200.SEG_DTYPE.itemsize == 16
    assert rpa_b200.RESULT_DTYPE.itemsize == 56


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import rpa_b200
    with pytest.raises(rpa_b200.TrpaError) as e:
        rpa_b200.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_band_geometry_and_schedule(tmp_path):
    """shapes.h on the host: the closed-form round gap is an upper bound of the exact requirement and the
    rotating schedule of the banded kernel never double-books a lane (tests/band_geometry_check.cpp)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "band_geometry_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(root, "taxator-tk_b200", "csrc"), "-o", exe,
                           os.path.join(root, "tests", "band_geometry_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok:")
