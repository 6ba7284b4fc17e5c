"""ctypes binding of the in-tree CUDA library libkzg.so (C ABI in include/kzg.h).

There is no CPU fallback: importing works anywhere, but every decode entry point needs the built
library and a CUDA device and raises loudly otherwise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_int, c_size_t, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KZG_LIB_PATH") or os.path.join(_HERE, "libkzg.so")      # (override: kernel variant sweeps, scripts/variant_sweep.sh)


class KzgError(RuntimeError):
    pass


class BatchStats(Structure):
    _fields_ = [("n_blocks", c_uint32), ("n_cblocks", c_uint32), ("n_hblocks", c_uint32), ("n_sblocks", c_uint32),
                ("lit_bytes", c_uint64), ("seq_records", c_uint64), ("workspace_bytes", c_uint64),
                ("launches", c_uint32), ("tile_bytes", c_uint32), ("sm_count", c_uint32), ("overflow", c_uint32)]


class InBuffer(Structure):
    _fields_ = [("src", c_void_p), ("size", c_size_t), ("pos", c_size_t)]


class OutBuffer(Structure):
    _fields_ = [("dst", c_void_p), ("size", c_size_t), ("pos", c_size_t)]


# every symbol include/kzg.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "kzg_isError", "kzg_getErrorName", "kzg_versionString", "kzg_ctx_create", "kzg_ctx_destroy", "kzg_scan_frames",
    "kzg_decompress_batch_device", "kzg_batch_finish", "kzg_decompress_batch_host", "kzg_host_alloc", "kzg_host_free",
    "kzg_last_batch_stats", "kzg_ctx_set_dictionary", "kzg_ctx_set_profiling", "kzg_last_stage_ms", "kzg_stage_ms_history", "kzg_createDCtx", "kzg_freeDCtx", "kzg_DCtx_loadDictionary", "kzg_decompressStream",
    "kzg_enc_create", "kzg_enc_destroy", "kzg_compress_bound", "kzg_compress_batch_device", "kzg_enc_finish",
    "kzg_compress_batch_host", "kzg_enc_set_profiling", "kzg_enc_last_stage_ms",
    "kzg_createCCtx", "kzg_freeCCtx", "kzg_CCtx_setParameter", "kzg_CCtx_loadDictionary", "kzg_compressStream2",
    "kzg_seek_table_bound", "kzg_write_seek_table",
    "kzg_segment_begin", "kzg_decompress_segment_host", "kzg_segment_checksum",
]

_lib = None


def lib() -> ctypes.CDLL:
    """Loads libkzg.so (never builds it: build() in __graft_entry__.py / kompressor_b200.build does)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KzgError(f"{LIB_PATH} is missing: run `python -m kompressor_b200.build` (there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    L.kzg_isError.argtypes = [c_size_t]
    L.kzg_isError.restype = ctypes.c_uint
    L.kzg_getErrorName.argtypes = [c_size_t]
    L.kzg_getErrorName.restype = c_char_p
    L.kzg_versionString.restype = c_char_p
    L.kzg_ctx_create.argtypes = [POINTER(c_void_p), c_int]
    L.kzg_ctx_create.restype = c_int
    L.kzg_ctx_destroy.argtypes = [c_void_p]
    L.kzg_ctx_destroy.restype = None
    L.kzg_scan_frames.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_size_t, POINTER(c_size_t)]
    L.kzg_scan_frames.restype = c_size_t
    L.kzg_decompress_batch_device.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_uint64, c_uint64, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p]
    L.kzg_decompress_batch_device.restype = c_int
    L.kzg_batch_finish.argtypes = [c_void_p]
    L.kzg_batch_finish.restype = c_int
    L.kzg_decompress_batch_host.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_uint64, c_void_p, c_void_p]
    L.kzg_decompress_batch_host.restype = c_int
    L.kzg_host_alloc.argtypes = [c_size_t]
    L.kzg_host_alloc.restype = c_void_p
    L.kzg_host_free.argtypes = [c_void_p]
    L.kzg_host_free.restype = None
    L.kzg_last_batch_stats.argtypes = [c_void_p, POINTER(BatchStats)]
    L.kzg_last_batch_stats.restype = c_int
    L.kzg_ctx_set_dictionary.argtypes = [c_void_p, c_void_p, c_size_t]
    L.kzg_ctx_set_dictionary.restype = c_int
    L.kzg_ctx_set_profiling.argtypes = [c_void_p, c_int]
    L.kzg_ctx_set_profiling.restype = c_int
    L.kzg_last_stage_ms.argtypes = [c_void_p, c_void_p, c_void_p, c_int]
    L.kzg_last_stage_ms.restype = c_int
    L.kzg_stage_ms_history.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int]
    L.kzg_stage_ms_history.restype = c_int
    L.kzg_enc_create.argtypes = [POINTER(c_void_p), c_int]
    L.kzg_enc_create.restype = c_int
    L.kzg_enc_destroy.argtypes = [c_void_p]
    L.kzg_enc_destroy.restype = None
    L.kzg_compress_bound.argtypes = [c_size_t]
    L.kzg_compress_bound.restype = c_size_t
    L.kzg_compress_batch_device.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64,
                                            c_uint64, c_int, c_void_p, c_void_p, c_void_p]
    L.kzg_compress_batch_device.restype = c_int
    L.kzg_enc_finish.argtypes = [c_void_p, c_void_p]
    L.kzg_enc_finish.restype = c_int
    L.kzg_compress_batch_host.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64,
                                          c_int, c_void_p, c_void_p]
    L.kzg_compress_batch_host.restype = c_int
    L.kzg_enc_set_profiling.argtypes = [c_void_p, c_int]
    L.kzg_enc_set_profiling.restype = c_int
    L.kzg_enc_last_stage_ms.argtypes = [c_void_p, c_void_p, c_void_p, c_int]
    L.kzg_enc_last_stage_ms.restype = c_int
    L.kzg_seek_table_bound.argtypes = [c_size_t]
    L.kzg_seek_table_bound.restype = c_size_t
    L.kzg_write_seek_table.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, c_size_t]
    L.kzg_write_seek_table.restype = c_size_t
    L.kzg_createCCtx.restype = c_void_p
    L.kzg_freeCCtx.argtypes = [c_void_p]
    L.kzg_freeCCtx.restype = c_size_t
    L.kzg_CCtx_setParameter.argtypes = [c_void_p, c_int, c_int]
    L.kzg_CCtx_setParameter.restype = c_size_t
    L.kzg_CCtx_loadDictionary.argtypes = [c_void_p, c_void_p, c_size_t]
    L.kzg_CCtx_loadDictionary.restype = c_size_t
    L.kzg_compressStream2.argtypes = [c_void_p, POINTER(OutBuffer), POINTER(InBuffer), c_int]
    L.kzg_compressStream2.restype = c_size_t
    if hasattr(L, "kzg_createDCtx"):
        L.kzg_createDCtx.restype = c_void_p
        L.kzg_freeDCtx.argtypes = [c_void_p]
        L.kzg_freeDCtx.restype = c_size_t
        L.kzg_DCtx_loadDictionary.argtypes = [c_void_p, c_void_p, c_size_t]
        L.kzg_DCtx_loadDictionary.restype = c_size_t
        L.kzg_decompressStream.argtypes = [c_void_p, POINTER(OutBuffer), POINTER(InBuffer)]
        L.kzg_decompressStream.restype = c_size_t
    _lib = L
    return L


def error_code(ret: int) -> int:
    """Positive ZSTD_ErrorCode of a size_t result (0 if it is not an error)."""
    ret &= (1 << 64) - 1
    return (1 << 64) - ret if ret > (1 << 64) - 120 else 0


def error_name(ret: int) -> str:
    return lib().kzg_getErrorName(ret & ((1 << 64) - 1)).decode()
