"""CPU-only: the C-ABI library loads and exports every symbol include/veritas_b200.h declares; without a GPU
the product fails loudly instead of falling back."""
import os
import re
import ctypes as C
import pytest

from veritas_b200._lib import load, SIGNATURES, LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "veritas_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vrt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB_PATH), "build with python -m veritas_b200.build"
    L = load()
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(L, s), s
        assert s in SIGNATURES, f"{s} has no ctypes signature"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = load()
    h = C.c_void_p()
    rc = L.vrt_create(C.byref(h), 0, 2)
    assert rc != 0
    assert b"no CPU fallback" in L.vrt_global_error()


def test_host_side_helpers_match_reference_formulas():
    L = load()
    # Settings::UpdateTime abscissae (Settings.cpp:166-179)
    t = 0.0
    for i in range(6):
        t = L.vrt_update_time(t, i, 1.0)
    assert abs(t - 1.0) < 1e-15
    # laser ramp is continuous and vanishes at t = 0
    assert L.vrt_case_laser_by(1e-6, 1.0, 0.0, 0.0) == 0.0
    assert L.vrt_case_maxwellian_slab(5e-6, 0.0, 3e-6, 7e-6, 2.0, 0.5) > 0
    assert L.vrt_case_maxwellian_slab(1e-6, 0.0, 3e-6, 7e-6, 2.0, 0.5) == 0.0
