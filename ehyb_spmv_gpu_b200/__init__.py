"""ehyb_spmv_gpu_b200 -- B200-native Explicit-Caching-HYB SpMV engine.

The engine itself is C + CUDA (csrc/, C ABI in include/*.h, built into lib/libehyb.so);
this package only loads it (no fallback) and mirrors the entry points for tests and bench.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "api"]
