"""kompressor_b200: B200-native Zstandard decompressor behind Kompressor's SliceTransform API.

The hot path is hand-written CUDA for sm_100a in csrc/ (libkzg.so, C ABI in include/kzg.h); this
package is the host-side mirror of the reference's operator interface for that path.
"""
from ._lib import KzgError, LIB_PATH, error_code, error_name  # noqa: F401

__all__ = ["KzgError", "LIB_PATH", "error_code", "error_name"]
