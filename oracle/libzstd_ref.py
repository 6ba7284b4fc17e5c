"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (kompressor_b200/).

ctypes harness around the *real* libzstd, i.e. the third-party library the
reference's hot path delegates to (``com.ensody.nativebuilds:zstd-libzstd:1.5.7.8``
= libzstd 1.5.7; reference: gradle/libs.versions.toml:9,46).  The Kotlin/JNI
wrapper itself cannot be built here (no JVM, no jni.h, no zstd.h), but the shared
library it binds is present in this image (Pillow bundles 1.5.7, the system has
1.5.5), so this module drives it with exactly the calls the wrapper makes:

* ``Wrapper.cpp:123-130``  ZSTD_createDCtx
* ``Wrapper.cpp:132-140``  ZSTD_freeDCtx
* ``Wrapper.cpp:58-73``    ZSTD_DCtx_loadDictionary
* ``Wrapper.cpp:142-187``  ZSTD_decompressStream(dctx, &out, &in) with absolute
  positions inside whole arrays
* ``Wrapper.cpp:189-196``  ZSTD_isError / ZSTD_getErrorName
* ``Wrapper.cpp:75-121``   ZSTD_compressStream2 (only to *produce* frames)

It is used (a) to pin the C restatement in ``zstd_oracle.c``, (b) to generate the
golden vectors under ``tests/golden`` and (c) as the correctness reference in the
GPU parity tests.  Nothing here reads ``/root/reference`` at run time.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import glob
import os
import sys
from ctypes import POINTER, Structure, byref, c_char_p, c_int, c_size_t, c_uint, c_ulonglong, c_void_p

ZSTD_CONTENTSIZE_UNKNOWN = (1 << 64) - 1
ZSTD_CONTENTSIZE_ERROR = (1 << 64) - 2

# ZSTD_cParameter / ZSTD_EndDirective values (public, stable ABI of zstd.h)
ZSTD_c_compressionLevel = 100
ZSTD_c_windowLog = 101
ZSTD_c_checksumFlag = 201
ZSTD_c_contentSizeFlag = 200
ZSTD_c_dictIDFlag = 202
ZSTD_e_continue, ZSTD_e_flush, ZSTD_e_end = 0, 1, 2
ZSTD_d_windowLogMax = 100


class ZSTD_inBuffer(Structure):
    _fields_ = [("src", c_void_p), ("size", c_size_t), ("pos", c_size_t)]


class ZSTD_outBuffer(Structure):
    _fields_ = [("dst", c_void_p), ("size", c_size_t), ("pos", c_size_t)]


def _candidates(prefer: str):
    site = [p for p in sys.path if p.endswith("site-packages")]
    pillow = []
    for s in site:
        pillow += sorted(glob.glob(os.path.join(s, "pillow.libs", "libzstd-*.so.1.5.7")))
    system = ["libzstd.so.1", "/usr/lib/x86_64-linux-gnu/libzstd.so.1"]
    if prefer == "1.5.7":
        return pillow + system
    return system + pillow


_CACHE: dict[str, "LibZstd"] = {}


class LibZstd:
    """One loaded libzstd.  ``prefer='1.5.7'`` picks the pinned version (frame generation,
    bit-exactness); ``prefer='fast'`` picks the system build, which decodes the same bytes
    faster and is what the CPU baseline is timed with."""

    def __init__(self, prefer: str = "1.5.7"):
        last = None
        self.lib = None
        for cand in _candidates(prefer):
            try:
                self.lib = ctypes.CDLL(cand)
                self.path = cand
                break
            except OSError as e:  # pragma: no cover
                last = e
        if self.lib is None:
            raise OSError(f"no libzstd found: {last}")
        L = self.lib
        L.ZSTD_versionString.restype = c_char_p
        L.ZSTD_versionNumber.restype = c_uint
        L.ZSTD_isError.argtypes = [c_size_t]
        L.ZSTD_isError.restype = c_uint
        L.ZSTD_getErrorName.argtypes = [c_size_t]
        L.ZSTD_getErrorName.restype = c_char_p
        L.ZSTD_createDCtx.restype = c_void_p
        L.ZSTD_freeDCtx.argtypes = [c_void_p]
        L.ZSTD_freeDCtx.restype = c_size_t
        L.ZSTD_DCtx_loadDictionary.argtypes = [c_void_p, c_void_p, c_size_t]
        L.ZSTD_DCtx_loadDictionary.restype = c_size_t
        L.ZSTD_DCtx_reset.argtypes = [c_void_p, c_int]
        L.ZSTD_DCtx_reset.restype = c_size_t
        L.ZSTD_decompressStream.argtypes = [c_void_p, POINTER(ZSTD_outBuffer), POINTER(ZSTD_inBuffer)]
        L.ZSTD_decompressStream.restype = c_size_t
        L.ZSTD_decompressDCtx.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t]
        L.ZSTD_decompressDCtx.restype = c_size_t
        L.ZSTD_decompress.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t]
        L.ZSTD_decompress.restype = c_size_t
        L.ZSTD_getFrameContentSize.argtypes = [c_void_p, c_size_t]
        L.ZSTD_getFrameContentSize.restype = c_ulonglong
        L.ZSTD_findFrameCompressedSize.argtypes = [c_void_p, c_size_t]
        L.ZSTD_findFrameCompressedSize.restype = c_size_t
        L.ZSTD_createCCtx.restype = c_void_p
        L.ZSTD_freeCCtx.argtypes = [c_void_p]
        L.ZSTD_freeCCtx.restype = c_size_t
        L.ZSTD_CCtx_setParameter.argtypes = [c_void_p, c_int, c_int]
        L.ZSTD_CCtx_setParameter.restype = c_size_t
        L.ZSTD_CCtx_setPledgedSrcSize.argtypes = [c_void_p, c_ulonglong]
        L.ZSTD_CCtx_setPledgedSrcSize.restype = c_size_t
        L.ZSTD_CCtx_loadDictionary.argtypes = [c_void_p, c_void_p, c_size_t]
        L.ZSTD_CCtx_loadDictionary.restype = c_size_t
        L.ZSTD_CCtx_reset.argtypes = [c_void_p, c_int]
        L.ZSTD_CCtx_reset.restype = c_size_t
        L.ZSTD_compressStream2.argtypes = [c_void_p, POINTER(ZSTD_outBuffer), POINTER(ZSTD_inBuffer), c_int]
        L.ZSTD_compressStream2.restype = c_size_t
        L.ZSTD_compress2.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t]
        L.ZSTD_compress2.restype = c_size_t
        L.ZSTD_compressBound.argtypes = [c_size_t]
        L.ZSTD_compressBound.restype = c_size_t
        if hasattr(L, "ZDICT_trainFromBuffer"):
            L.ZDICT_trainFromBuffer.argtypes = [c_void_p, c_size_t, c_void_p, POINTER(c_size_t), c_uint]
            L.ZDICT_trainFromBuffer.restype = c_size_t
        self.version = L.ZSTD_versionString().decode()

    # ------------------------------------------------------------------ dictionaries
    def train_dictionary(self, samples: list[bytes], capacity: int = 4096) -> bytes:
        """ZDICT_trainFromBuffer: a *formatted* dictionary (magic 0xEC30A437: dictID, entropy tables, repeat
        offsets, content) trained on `samples`."""
        import ctypes
        blob = b"".join(samples)
        sizes = (c_size_t * len(samples))(*[len(x) for x in samples])
        out = ctypes.create_string_buffer(capacity)
        n = self.lib.ZDICT_trainFromBuffer(out, capacity, blob, sizes, len(samples))
        if self.is_error(n):
            raise RuntimeError("ZDICT_trainFromBuffer: " + self.error_name(n))
        return out.raw[:n]

    # ------------------------------------------------------------------ errors
    def is_error(self, code: int) -> bool:
        return bool(self.lib.ZSTD_isError(code))

    def error_name(self, code: int) -> str:
        return self.lib.ZSTD_getErrorName(code & ((1 << 64) - 1)).decode()

    @staticmethod
    def error_code(code: int) -> int:
        """Positive ZSTD_ErrorCode of an error return value (0 if not an error)."""
        code &= (1 << 64) - 1
        return (1 << 64) - code if code > (1 << 64) - 120 else 0

    # ------------------------------------------------------------- compression
    def compress(self, data: bytes, level: int = 3, checksum: bool = False, pledged: bool = True,
                 dictionary: bytes | None = None, window_log: int | None = None,
                 chunk: int | None = None) -> bytes:
        """One frame.  ``pledged=True`` is the one-shot ``ZSTD_e_end`` call
        (``SliceTransform.transform(ByteArray)`` shape, single-segment header with FCS);
        ``pledged=False`` streams ``chunk``-byte ``ZSTD_e_continue`` pieces the way the
        ``pipe()`` path does (window descriptor, no FCS)."""
        L = self.lib
        cctx = L.ZSTD_createCCtx()
        try:
            self._chk(L.ZSTD_CCtx_setParameter(cctx, ZSTD_c_compressionLevel, level))
            if checksum:
                self._chk(L.ZSTD_CCtx_setParameter(cctx, ZSTD_c_checksumFlag, 1))
            if window_log is not None:
                self._chk(L.ZSTD_CCtx_setParameter(cctx, ZSTD_c_windowLog, window_log))
            if dictionary is not None:
                self._chk(L.ZSTD_CCtx_loadDictionary(cctx, dictionary, len(dictionary)))
            bound = L.ZSTD_compressBound(len(data)) + 64
            dst = ctypes.create_string_buffer(bound)
            if pledged:
                n = L.ZSTD_compress2(cctx, dst, bound, data, len(data))
                self._chk(n)
                return dst.raw[:n]
            src = ctypes.create_string_buffer(data, len(data)) if data else ctypes.create_string_buffer(1)
            out = ZSTD_outBuffer(ctypes.cast(dst, c_void_p), bound, 0)
            chunk = chunk or 8192
            pos = 0
            while pos < len(data):
                end = min(len(data), pos + chunk)
                inb = ZSTD_inBuffer(ctypes.cast(src, c_void_p), end, pos)
                while inb.pos < inb.size:
                    self._chk(L.ZSTD_compressStream2(cctx, byref(out), byref(inb), ZSTD_e_continue))
                pos = end
            inb = ZSTD_inBuffer(ctypes.cast(src, c_void_p), len(data), len(data))
            while True:
                r = L.ZSTD_compressStream2(cctx, byref(out), byref(inb), ZSTD_e_end)
                self._chk(r)
                if r == 0:
                    break
            return dst.raw[:out.pos]
        finally:
            L.ZSTD_freeCCtx(cctx)

    def _chk(self, code: int) -> int:
        if self.lib.ZSTD_isError(code):
            raise RuntimeError(f"zstd error {self.error_code(code)}: {self.error_name(code)}")
        return code

    # ----------------------------------------------------------- decompression
    def decompress_oneshot(self, frame: bytes, capacity: int, dictionary: bytes | None = None):
        """ZSTD_decompressDCtx on (possibly concatenated) frames -> (bytes|None, error_code)."""
        L = self.lib
        dctx = L.ZSTD_createDCtx()
        try:
            if dictionary is not None:
                self._chk(L.ZSTD_DCtx_loadDictionary(dctx, dictionary, len(dictionary)))
            dst = ctypes.create_string_buffer(max(capacity, 1))
            n = L.ZSTD_decompressDCtx(dctx, dst, capacity, frame, len(frame))
            if L.ZSTD_isError(n):
                return None, self.error_code(n)
            return dst.raw[:n], 0
        finally:
            L.ZSTD_freeDCtx(dctx)

    def decompress_stream(self, data: bytes, dictionary: bytes | None = None,
                          out_slice: int | None = None):
        """The reference's exact call pattern for ``transform(ByteArray)``
        (SliceTransform.kt:33-45 -> ZstdDecompressor.jvm.kt:25-39 -> Wrapper.cpp:142-187):
        whole input visible, output slices of ``max(8192, n/10)`` bytes, a new slice whenever
        ``output.isFull && result != 0``.  Returns (output bytes, error_code, n_calls)."""
        L = self.lib
        dctx = L.ZSTD_createDCtx()
        try:
            if dictionary is not None:
                r = L.ZSTD_DCtx_loadDictionary(dctx, dictionary, len(dictionary))
                if L.ZSTD_isError(r):
                    return b"", self.error_code(r), 0
            n = len(data)
            block = out_slice or max(8192, n // 10)
            src = ctypes.create_string_buffer(data, n) if n else ctypes.create_string_buffer(1)
            inb = ZSTD_inBuffer(ctypes.cast(src, c_void_p), n, 0)
            pieces = []
            calls = 0
            while True:
                dst = ctypes.create_string_buffer(block)
                out = ZSTD_outBuffer(ctypes.cast(dst, c_void_p), block, 0)
                insufficient = False
                r = L.ZSTD_decompressStream(dctx, byref(out), byref(inb))
                calls += 1
                if L.ZSTD_isError(r):
                    pieces.append(dst.raw[:out.pos])
                    return b"".join(pieces), self.error_code(r), calls
                insufficient = (out.pos == out.size) and r != 0
                pieces.append(dst.raw[:out.pos])
                if not (inb.pos != inb.size or insufficient):
                    break
            return b"".join(pieces), 0, calls
        finally:
            L.ZSTD_freeDCtx(dctx)

    def frame_content_size(self, frame: bytes) -> int:
        return self.lib.ZSTD_getFrameContentSize(frame, len(frame))

    def frame_compressed_size(self, data: bytes) -> int:
        return self._chk(self.lib.ZSTD_findFrameCompressedSize(data, len(data)))


def load(prefer: str = "1.5.7") -> LibZstd:
    if prefer not in _CACHE:
        _CACHE[prefer] = LibZstd(prefer)
    return _CACHE[prefer]


if __name__ == "__main__":
    z = load()
    print(z.path, z.version)
    f = z.compress(b"hello compression world" * 100, 3, checksum=True)
    print(len(f), z.decompress_stream(f)[0][:23])
