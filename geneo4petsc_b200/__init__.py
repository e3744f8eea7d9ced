"""geneo4petsc_b200 -- B200-native (sm_100a) GenEO Schwarz preconditioner + Krylov hot path of geneo4PETSc.

The product is the C ABI shared library ``libgeneob200.so`` (C++ host + hand-written CUDA, include/geneo_b200.h).
This Python package is only the thin ctypes mirror used by tests/, bench.py and __graft_entry__.py.
It never imports anything under oracle/ and has no CPU fallback: numeric calls raise when no CUDA device is present.
"""
from .api import (GeneoError, Problem, GeneoPC, Symbolic, lib, device_count, host_sym_eig, microbench, counters, profile_dump,  # noqa: F401
                  KSP_REASONS)
