"""integrals_b200 -- B200-native batched quadrature engine behind the Integrals.jl solve interface.

Layout: csrc/ (CUDA kernels + the C ABI of include/integrals_b200.h -> libintegrals_b200.so),
_lib.py (ctypes binding), and the host-side mirror of the reference's problem / algorithm / solve
interface.  There is no CPU fallback: without the built library and a B200 the solve calls raise.
"""
from ._lib import (Engine, IB200Error, FAMILY_IDS, TRANSFORM_IDS, RULE_IDS, SYMBOLS, LIB_PATH,
                   MAXEVALS_UNBOUNDED, default_engine, load_library)
from .build import build_library

__all__ = ["Engine", "IB200Error", "FAMILY_IDS", "TRANSFORM_IDS", "RULE_IDS", "SYMBOLS", "LIB_PATH",
           "MAXEVALS_UNBOUNDED", "default_engine", "load_library", "build_library"]
