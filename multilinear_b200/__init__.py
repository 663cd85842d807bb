"""multilinear_b200 — B200 (sm_100a) CUDA backend for the polynomial-commitment hot path of fr34za/multilinear.

The product is libmultilinear_b200.so (C ABI in include/multilinear_b200.h); this package is the thin
Python host-side mirror of the reference's interface used by the tests and bench.py.
"""
from . import api  # noqa: F401
from ._lib import MlError, NotPowerOfTwo, NotRsCode, SizeMismatch, lib_path, load  # noqa: F401
from .build import build  # noqa: F401
