"""helib_b200 -- B200-native DoubleCRT / NTT / key-switch engine behind HElib's API.

The product is the C-ABI shared library helib_b200/libhelib_b200.so (include/helib_b200.h),
built from helib_b200/csrc/*.cu for sm_100a.  This Python package is plumbing only: a ctypes
binding used by the tests, bench.py and __graft_entry__.py.  There is no CPU fallback: importing
works anywhere, but creating an Engine without a CUDA device (or without the built library)
raises.
"""
from .engine import Engine, Poly, Chain, HbError, load_library, library_path  # noqa: F401
from .build import build_library  # noqa: F401

__all__ = ["Engine", "Poly", "Chain", "HbError", "load_library", "library_path", "build_library"]
