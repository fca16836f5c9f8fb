"""veritas_b200 — B200-native (sm_100a) implementation of the Veritas 1D1P Vlasov advance.

The product is the C-ABI shared library `libveritas_b200.so` (include/veritas_b200.h) and the C++ host
classes in veritas_b200/host/.  This Python package is only the ctypes front-end used by tests and bench.py.
There is no CPU fallback: `load()` raises if the library is missing, and `Context()` raises without a GPU.
"""
from ._lib import load, VrtError, PatchDesc, CaseParams, CaseDerived  # noqa: F401
from .solver import Context, LaserPlasmaRun  # noqa: F401
