"""TEST INFRASTRUCTURE: ctypes front-end of tests/hostsim/hostsim.cpp (the CUDA decoder's per-thread
stage code compiled for the host).  Never used by the product path."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhostsim.so")
_CSRC = os.path.join(_HERE, "..", "..", "kompressor_b200", "csrc")


def build():
    deps = [os.path.join(_HERE, "hostsim.cpp"), os.path.join(_HERE, "simt_emu.h")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        subprocess.check_call(["g++", "-O1", "-fPIC", "-shared", "-std=c++17", "-x", "c++", "-Wall", "-Wno-unknown-pragmas",
                               "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-I", _HERE, "-o", _SO, os.path.join(_HERE, "hostsim.cpp")])
    return _SO


_lib = None


def run(frames, caps, exec_mode: int = 0, tile: int = 8192):
    """Decodes `frames` (list of bytes) with the host-compiled stage code -> (outputs, status, sizes).
    exec_mode 0: sequential execute stand-in; 1 / 2: the second execute design (kzg_exec2.cuh) on the lock-step warp
    emulator with lanes scheduled in forward / reverse order, ring size `tile`."""
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.kzs_decompress_batch.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                              ctypes.c_void_p, ctypes.c_uint64]
    _lib.kzs_set_exec.argtypes = [ctypes.c_int, ctypes.c_uint32]
    _lib.kzs_set_exec.restype = None
    _lib.kzs_exec_deferred.restype = ctypes.c_uint64
    _lib.kzs_set_exec(exec_mode, tile)
    n = len(frames)
    blob = b"".join(frames)
    comp = np.frombuffer(blob, dtype=np.uint8).copy() if blob else np.zeros(1, np.uint8)
    sizes = np.array([len(f) for f in frames], dtype=np.uint64)
    offs = np.zeros(n, np.uint64)
    offs[1:] = np.cumsum(sizes)[:-1]
    caps = np.array(caps, dtype=np.uint64)
    doffs = np.zeros(n, np.uint64)
    doffs[1:] = np.cumsum(caps)[:-1]
    dst = np.zeros(int(caps.sum()) + 64, np.uint8)
    status = np.zeros(n, np.uint32)
    osz = np.zeros(n, np.uint64)
    r = _lib.kzs_decompress_batch(comp.ctypes.data, int(sizes.sum()), offs.ctypes.data, sizes.ctypes.data, n,
                                  dst.ctypes.data, doffs.ctypes.data, caps.ctypes.data, status.ctypes.data,
                                  osz.ctypes.data, None, 0, None, 0)
    assert r == 0, r
    return [bytes(dst[int(doffs[i]):int(doffs[i] + osz[i])]) for i in range(n)], status, osz


def compress(chunks, checksum=False):
    """Compresses `chunks` (list of bytes, each <= 128 KiB) with the host-compiled encode stage behind a sequential
    matcher -> (frames, status)."""
    global _lib
    if _lib is None:
        run([], [])
    _lib.kzs_compress_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    n = len(chunks)
    blob = b"".join(chunks)
    src = np.frombuffer(blob + b"\0" * 16, dtype=np.uint8).copy()
    sizes = np.array([len(c) for c in chunks], dtype=np.uint64)
    offs = np.zeros(n, np.uint64)
    offs[1:] = np.cumsum(sizes)[:-1]
    caps = sizes + np.uint64(32)
    doffs = np.zeros(n, np.uint64)
    doffs[1:] = np.cumsum(caps)[:-1]
    dst = np.zeros(int(caps.sum()) + 64, np.uint8)
    status = np.zeros(n, np.uint32)
    osz = np.zeros(n, np.uint64)
    r = _lib.kzs_compress_batch(src.ctypes.data, offs.ctypes.data, sizes.ctypes.data, n, dst.ctypes.data, doffs.ctypes.data,
                                caps.ctypes.data, 1 if checksum else 0, status.ctypes.data, osz.ctypes.data)
    assert r == 0, r
    return [bytes(dst[int(doffs[i]):int(doffs[i] + osz[i])]) for i in range(n)], status


def weights_roundtrip(weights) -> int:
    """Encoder's FSE-coded Huffman weight description -> decoder's parser. Returns the size (0: declined), -1: mismatch."""
    global _lib
    if _lib is None:
        run([], [])
    w = np.asarray(weights, dtype=np.uint8)
    _lib.kzs_weights_roundtrip.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
    return int(_lib.kzs_weights_roundtrip(w.ctypes.data, w.size))
