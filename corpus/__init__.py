"""BENCH/TEST INFRASTRUCTURE: synthetic corpora of BASELINE.json's configs.

`build(...)` generates frames with the C generator in corpus.c and compresses each one
independently with the real libzstd 1.5.7 (the frames the reference's ZstdCompressor would
produce).  Not part of the product path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkzgcorpus.so")
KINDS = {"logs": 0, "json": 1, "literal": 2, "random": 3, "mixed": 4}
SEED = 20261018


def build_lib(force: bool = False) -> str:
    src = os.path.join(_HERE, "corpus.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-o", _SO, src, "-ldl", "-lpthread", "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build_lib())
        L.kzc_gen.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
        L.kzc_gen.restype = None
        L.kzc_frame_seed.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        L.kzc_frame_seed.restype = ctypes.c_uint64
        L.kzc_slot_size.argtypes = [ctypes.c_size_t]
        L.kzc_slot_size.restype = ctypes.c_size_t
        L.kzc_build.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.kzc_build.restype = ctypes.c_uint64
        _lib = L
    return _lib


def libzstd_path(prefer: str = "1.5.7") -> str:
    import glob
    import sys
    cands = []
    for s in sys.path:
        if s.endswith("site-packages"):
            cands += sorted(glob.glob(os.path.join(s, "pillow.libs", "libzstd-*.so.1.5.7")))
    system = [p for p in ("/usr/lib/x86_64-linux-gnu/libzstd.so.1", "/lib/x86_64-linux-gnu/libzstd.so.1") if os.path.exists(p)]
    order = cands + system if prefer == "1.5.7" else system + cands
    if not order:
        raise OSError("no libzstd shared library found")
    return order[0]


def generate(kind: str, seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    lib().kzc_gen(KINDS[kind], seed & (2**64 - 1), out.ctypes.data, n)
    return out


@dataclass
class Corpus:
    kind: str
    frame_size: int
    n_frames: int
    comp: np.ndarray      # concatenated compressed frames (uint8)
    offsets: np.ndarray   # uint64 [n_frames] start of each frame in comp
    sizes: np.ndarray     # uint64 [n_frames] compressed size
    raw: np.ndarray | None  # uint8 [n_frames*frame_size] expected output, if kept
    level: int
    checksum: bool
    streamed: bool

    @property
    def raw_bytes(self) -> int:
        return self.n_frames * self.frame_size

    @property
    def comp_bytes(self) -> int:
        return int(self.comp.size)

    @property
    def ratio(self) -> float:
        return self.raw_bytes / max(1, self.comp_bytes)


def build(kind: str, n_frames: int, frame_size: int = 65536, level: int = 3, checksum: bool = False,
          streamed: bool = False, seed: int = SEED, first_frame: int = 0, keep_raw: bool = True,
          threads: int | None = None) -> Corpus:
    L = lib()
    threads = threads or min(os.cpu_count() or 1, 64)
    slot = L.kzc_slot_size(frame_size)
    comp = np.empty(n_frames * slot, dtype=np.uint8)
    sizes = np.zeros(n_frames, dtype=np.uint64)
    raw = np.empty(n_frames * frame_size, dtype=np.uint8) if keep_raw else None
    total = L.kzc_build(libzstd_path("1.5.7").encode(), KINDS[kind], seed, first_frame, n_frames, frame_size, level,
                        int(checksum), int(streamed), threads, raw.ctypes.data if raw is not None else None,
                        comp.ctypes.data, sizes.ctypes.data)
    if total == 0 and n_frames:
        raise RuntimeError("corpus build failed")
    comp = comp[:total].copy()
    offsets = np.zeros(n_frames, dtype=np.uint64)
    if n_frames:
        offsets[1:] = np.cumsum(sizes)[:-1]
    return Corpus(kind, frame_size, n_frames, comp, offsets, sizes, raw, level, checksum, streamed)
