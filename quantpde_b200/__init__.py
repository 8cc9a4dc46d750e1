"""quantpde_b200 — B200-native batched time-stepping sweep with the call shape of
QuantPDE's stepper / discretisation / penalty / solver classes.

Product path = quantpde_b200/lib/libqpde_b200.so (hand-written sm_100a CUDA,
C ABI in include/qpde_b200.h) + the header-only C++ host API in
quantpde_b200/include/QuantPDE_B200.  This Python package is only the ctypes
plumbing that tests/ and bench.py use to reach the C ABI.
"""
from . import capi  # noqa: F401
