"""TEST INFRASTRUCTURE ONLY.  ctypes front-end of the C oracle (zstd_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product package (kompressor_b200) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
BASELINE_SO = os.path.join(_HERE, "_ref", "libkzg_cpu_baseline.so")


def build(force: bool = False) -> None:
    """Compiles liboracle.so and oracle/_ref/libkzg_cpu_baseline.so with the committed Makefile."""
    srcs = [os.path.join(_HERE, f) for f in ("zstd_oracle.c", "zstd_oracle.h", "cpu_baseline.c", "Makefile")]
    newest = max(os.path.getmtime(s) for s in srcs if os.path.exists(s))
    targets = [_SO, BASELINE_SO]
    if force or any(not os.path.exists(t) or os.path.getmtime(t) < newest for t in targets):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-f", os.path.join(_HERE, "Makefile")])


class BlockInfo(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("last", ctypes.c_int), ("content_size", ctypes.c_uint32),
                ("regen_size", ctypes.c_uint32), ("lit_type", ctypes.c_int), ("lit_size", ctypes.c_uint32),
                ("n_streams", ctypes.c_int), ("n_seq", ctypes.c_uint32), ("modes", ctypes.c_int * 3),
                ("logs", ctypes.c_int * 3), ("huf_maxbits", ctypes.c_int)]


class Seq(ctypes.Structure):
    _fields_ = [("ll", ctypes.c_uint32), ("ml", ctypes.c_uint32), ("offset", ctypes.c_uint32)]


class FrameHeader(ctypes.Structure):
    _fields_ = [("window_size", ctypes.c_uint64), ("content_size", ctypes.c_uint64), ("dict_id", ctypes.c_uint32),
                ("single_segment", ctypes.c_int), ("has_checksum", ctypes.c_int), ("header_size", ctypes.c_uint32)]


_BLOCK_CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.POINTER(BlockInfo), ctypes.POINTER(ctypes.c_uint8),
                             ctypes.POINTER(Seq))
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.kzo_decompress.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                     ctypes.c_char_p, ctypes.c_size_t]
        L.kzo_decompress.restype = ctypes.c_size_t
        L.kzo_decompress_ex.argtypes = L.kzo_decompress.argtypes + [_BLOCK_CB, ctypes.c_void_p]
        L.kzo_decompress_ex.restype = ctypes.c_size_t
        L.kzo_frame_header_parse.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(FrameHeader)]
        L.kzo_frame_header_parse.restype = ctypes.c_size_t
        L.kzo_find_frame_size.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.kzo_find_frame_size.restype = ctypes.c_size_t
        L.kzo_xxh64.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64]
        L.kzo_xxh64.restype = ctypes.c_uint64
        L.kzo_is_error.argtypes = [ctypes.c_size_t]
        L.kzo_error_name.argtypes = [ctypes.c_size_t]
        L.kzo_error_name.restype = ctypes.c_char_p
        _lib = L
    return _lib


def error_code(ret: int) -> int:
    ret &= (1 << 64) - 1
    return (1 << 64) - ret if ret > (1 << 64) - 120 else 0


def decompress(data: bytes, capacity: int, dictionary: bytes | None = None):
    """-> (bytes | None, error_code)"""
    dst = ctypes.create_string_buffer(max(1, capacity))
    r = lib().kzo_decompress(data, len(data), dst, capacity, dictionary, len(dictionary) if dictionary else 0)
    e = error_code(r)
    if e:
        return None, e
    return dst.raw[:r], 0


def decompress_blocks(data: bytes, capacity: int, dictionary: bytes | None = None):
    """-> (bytes | None, error_code, [ {info..., 'literals': bytes, 'seqs': [(ll, ml, off)]} ])"""
    blocks = []

    def cb(_user, info, lit, seqs):
        i = info.contents
        d = {k: getattr(i, k) for k, _ in BlockInfo._fields_ if k not in ("modes", "logs")}
        d["modes"] = list(i.modes)
        d["logs"] = list(i.logs)
        d["literals"] = bytes(lit[:i.lit_size]) if lit and i.type == 2 else b""
        d["seqs"] = [(seqs[k].ll, seqs[k].ml, seqs[k].offset) for k in range(i.n_seq)] if seqs and i.type == 2 else []
        blocks.append(d)

    dst = ctypes.create_string_buffer(max(1, capacity))
    r = lib().kzo_decompress_ex(data, len(data), dst, capacity, dictionary, len(dictionary) if dictionary else 0,
                                _BLOCK_CB(cb), None)
    e = error_code(r)
    if e:
        return None, e, blocks
    return dst.raw[:r], 0, blocks


def made_up_literals() -> bool:
    """True if a frame decoded since the last call ended for the one reason documented as a divergence from libzstd
    (a Huffman stream ran out of bits where libzstd's fast loops make up literals); clears the flag."""
    return bool(lib().kzo_made_up_literals())


def xxh64(data: bytes, seed: int = 0) -> int:
    return lib().kzo_xxh64(data, len(data), seed)


def find_frame_size(data: bytes) -> int:
    return lib().kzo_find_frame_size(data, len(data))


def frame_header(data: bytes):
    h = FrameHeader()
    r = lib().kzo_frame_header_parse(data, len(data), ctypes.byref(h))
    return h, error_code(r)


def cpu_baseline(corpus, threads: int, passes: int = 3, mode: int = 0, prefer: str = "fast"):
    """Times the real libzstd on `corpus` (a corpus.Corpus) on the host cores.
    mode 0 = parallel one-shot, mode 1 = reference call pattern (see cpu_baseline.c).
    Returns dict(gbps=best-pass decompressed GB/s, seconds=[...], version, path, threads, bytes)."""
    import numpy as np
    from corpus import libzstd_path
    if not os.path.exists(BASELINE_SO):
        build()
    L = ctypes.CDLL(BASELINE_SO)
    L.kzb_run.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                          ctypes.c_uint64, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                          ctypes.c_char_p]
    L.kzb_run.restype = ctypes.c_uint64
    path = libzstd_path(prefer)
    t = np.zeros(passes, dtype=np.float64)
    ver = ctypes.create_string_buffer(32)
    offs = np.ascontiguousarray(corpus.offsets, dtype=np.uint64)
    sizes = np.ascontiguousarray(corpus.sizes, dtype=np.uint64)
    total = L.kzb_run(path.encode(), mode, corpus.comp.ctypes.data, offs.ctypes.data, sizes.ctypes.data,
                      corpus.n_frames, corpus.frame_size, threads, passes, t.ctypes.data, ver)
    if total == 0:
        raise RuntimeError("cpu baseline failed")
    timed = t[1:] if passes > 1 else t
    return {"gbps": total / float(timed.min()) / 1e9, "seconds": [float(x) for x in t], "version": ver.value.decode(),
            "path": path, "threads": threads, "bytes": int(total), "mode": mode}
