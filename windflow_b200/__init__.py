"""windflow_b200 -- B200-native (sm_100a) kernels for the WindFlow GPU operator hot path.

The product is libwfb200.so (C ABI in include/wfb200.h). This package is the Python-side plumbing used by the
tests and the benchmark; importing it does not load the library, using it does -- and fails loudly when the
library has not been built (there is no CPU fallback).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops", "build"]
