"""B200-native KZG engine — drop-in for the hot path of VGLoic/kzg-poly-commit-exploration.

The product is libkzg_b200.so (hand-written sm_100a CUDA behind the C ABI in include/kzg_b200.h).
This package holds the ctypes binding and a host-side mirror of the reference's public API
(`scalar`, `curves`, `polynomial`, `trusted_setup` — same names as the Rust modules)."""
from ._native import KzgNativeError, LIB_PATH, lib  # noqa: F401
from .engine import Context, Setup  # noqa: F401
