"""Host-side mirror of kompressor-zstd's decompressor binding, on top of the kzg C ABI.

Reference:
  ZstdDecompressor(dictionary)   kompressor-zstd--nativelib/src/nonJsMain/.../Zstd.nonjs.kt:11
  ZstdDecompressorImpl           kompressor-zstd--nativelib/src/jvmCommonMain/kotlin/com/ensody/kompressor/zstd/ZstdDecompressor.jvm.kt:10-40
  ZstdWrapper.decompressStream   .../ZstdWrapper.kt:48-58  ->  Wrapper.cpp:142-187
  checkErrorResult               .../ZstdCompressor.jvm.kt:45-51
"""
from __future__ import annotations

import ctypes

from .._lib import InBuffer, KzgError, OutBuffer, lib
from .core import ByteArraySlice, SliceTransform


class IllegalStateException(RuntimeError):
    """What Kotlin's error()/check() throw; `code` is the positive ZSTD_ErrorCode when there is one."""

    def __init__(self, message: str, code: int = 0):
        super().__init__(message)
        self.code = code


def checkErrorResult(result: int) -> None:
    """ZstdCompressor.jvm.kt:45-51: throws if `result` is a zstd error. `result` prints as a JVM Long."""
    if result != 0:
        L = lib()
        if L.kzg_isError(result & ((1 << 64) - 1)):
            name = L.kzg_getErrorName(result & ((1 << 64) - 1)).decode()
            as_long = result - (1 << 64) if result >= (1 << 63) else result
            raise IllegalStateException(f"Bad zstd result code {as_long}: {name}", code=-as_long)


class ZstdDecompressorImpl(SliceTransform):
    def __init__(self, dictionary: bytes | None = None):
        self._L = lib()
        self.dctx = self._L.kzg_createDCtx()
        if not self.dctx:
            raise IllegalStateException("Failed allocating zstd dctx (kzg needs a CUDA device; there is no CPU fallback)")
        self.dictionary = dictionary
        if dictionary is not None:
            buf = (ctypes.c_char * len(dictionary)).from_buffer_copy(dictionary)
            checkErrorResult(self._L.kzg_DCtx_loadDictionary(self.dctx, buf, len(dictionary)))

    def close(self) -> None:                      # the JVM does this from the Cleaner thread
        if getattr(self, "dctx", None):
            self._L.kzg_freeDCtx(self.dctx)
            self.dctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform(self, input: ByteArraySlice, output: ByteArraySlice, finish: bool) -> None:  # noqa: A002
        # Wrapper.cpp:167-176: buffers cover the whole arrays, positions are absolute
        in_arr = (ctypes.c_char * max(1, len(input.data))).from_buffer(input.data) if len(input.data) else None
        out_arr = (ctypes.c_char * max(1, len(output.data))).from_buffer(output.data) if len(output.data) else None
        inb = InBuffer(ctypes.addressof(in_arr) if in_arr is not None else None, input.writeStart, input.readStart)
        outb = OutBuffer(ctypes.addressof(out_arr) if out_arr is not None else None, output.writeLimit, output.writeStart)
        result = self._L.kzg_decompressStream(self.dctx, ctypes.byref(outb), ctypes.byref(inb))
        input.readStart = inb.pos
        output.writeStart = outb.pos
        del in_arr, out_arr
        checkErrorResult(result)
        output.insufficient = output.isFull and result != 0


class ZstdCompressorImpl(SliceTransform):
    """ZstdCompressor.jvm.kt:10-42 on top of the kzg frame writer (compress side of the shim).  The stream is written
    as independent 64 KiB frames; `compressionLevel` below 3 selects the greedy parse, from 3 one step of lazy matching."""

    def __init__(self, compressionLevel: int = 3, dictionary: bytes | None = None):
        self._L = lib()
        self.cctx = self._L.kzg_createCCtx()
        if not self.cctx:
            raise IllegalStateException("Failed allocating zstd cctx (kzg needs a CUDA device; there is no CPU fallback)")
        self._L.kzg_CCtx_setParameter(self.cctx, 100, compressionLevel)        # ZstdParameter.compressionLevel
        if dictionary is not None:
            buf = (ctypes.c_char * len(dictionary)).from_buffer_copy(dictionary)
            checkErrorResult(self._L.kzg_CCtx_loadDictionary(self.cctx, buf, len(dictionary)))

    def setParameter(self, parameter: int, value: int) -> None:                # ZstdWrapper.setParameter
        checkErrorResult(self._L.kzg_CCtx_setParameter(self.cctx, parameter, value))

    def close(self) -> None:
        if getattr(self, "cctx", None):
            self._L.kzg_freeCCtx(self.cctx)
            self.cctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform(self, input: ByteArraySlice, output: ByteArraySlice, finish: bool) -> None:  # noqa: A002
        in_arr = (ctypes.c_char * max(1, len(input.data))).from_buffer(input.data) if len(input.data) else None
        out_arr = (ctypes.c_char * max(1, len(output.data))).from_buffer(output.data) if len(output.data) else None
        inb = InBuffer(ctypes.addressof(in_arr) if in_arr is not None else None, input.writeStart, input.readStart)
        outb = OutBuffer(ctypes.addressof(out_arr) if out_arr is not None else None, output.writeLimit, output.writeStart)
        result = self._L.kzg_compressStream2(self.cctx, ctypes.byref(outb), ctypes.byref(inb), 2 if finish else 0)
        input.readStart = inb.pos
        output.writeStart = outb.pos
        del in_arr, out_arr
        checkErrorResult(result)
        output.insufficient = input.hasData or (finish and result != 0)        # ZstdCompressor.jvm.kt:40


def ZstdCompressor(compressionLevel: int = 3, dictionary: bytes | None = None) -> SliceTransform:
    """public expect fun ZstdCompressor(compressionLevel: Int = 3, dictionary: ByteArray? = null): SliceTransform"""
    return ZstdCompressorImpl(compressionLevel=compressionLevel, dictionary=dictionary)


def ZstdDecompressor(dictionary: bytes | None = None) -> SliceTransform:
    """public expect fun ZstdDecompressor(dictionary: ByteArray? = null): SliceTransform"""
    return ZstdDecompressorImpl(dictionary=dictionary)


__all__ = ["ZstdCompressor", "ZstdCompressorImpl", "ZstdDecompressor", "ZstdDecompressorImpl", "IllegalStateException", "checkErrorResult", "KzgError"]
