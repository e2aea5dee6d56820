"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/mfb.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from matfac_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "mfb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mfb_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(E.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = E.load_library()
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(E.EngineError) as ei:
        E.Engine(10, 10, 4)
    assert "no CUDA device" in str(ei.value)


def test_argument_validation_without_device():
    lib = E.load_library()
    h = ctypes.c_void_p()
    cfg = E.Config(0, 0, 10, 4)
    assert lib.mfb_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"empty matrix" in lib.mfb_last_error()
    cfg = E.Config(0, 10, 10, 1000)
    assert lib.mfb_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"rank" in lib.mfb_last_error()


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "matfac_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "mf_oracle" not in txt and "oracle_lib" not in txt and "oracle/" not in txt, os.path.join(dp, f)
