"""Layer A: batched decode of independent zstd frames on one B200 (include/kzg.h, kzg_decompress_batch_*).

torch is used only as the owner of device memory and streams; all work happens in libkzg.so.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from ._lib import BatchStats, KzgError, lib


def scan_frames(buf) -> tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """Host-side frame table of a buffer of concatenated frames -> (offsets, sizes, content_sizes, consumed)."""
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    L = lib()
    cap = max(16, a.size // 9 + 1)
    off = np.zeros(cap, np.uint64)
    size = np.zeros(cap, np.uint64)
    content = np.zeros(cap, np.uint64)
    consumed = ctypes.c_size_t(0)
    n = L.kzg_scan_frames(a.ctypes.data if a.size else None, a.size, off.ctypes.data, size.ctypes.data,
                          content.ctypes.data, cap, ctypes.byref(consumed))
    if L.kzg_isError(n):
        raise KzgError(f"zstd error {(1 << 64) - n}: {L.kzg_getErrorName(n).decode()}")
    return off[:n].copy(), size[:n].copy(), content[:n].copy(), consumed.value


@dataclass
class DeviceBatch:
    """A batch resident in HBM: torch uint8/int64 tensors on one CUDA device."""
    src: "torch.Tensor"
    src_off: "torch.Tensor"
    src_size: "torch.Tensor"
    dst: "torch.Tensor"
    dst_off: "torch.Tensor"
    dst_cap: "torch.Tensor"
    status: "torch.Tensor"
    out_size: "torch.Tensor"
    n_frames: int
    total_src: int
    total_dst: int
    max_dst: int


class BatchDecoder:
    """One decode context (workspace + kernels) on one GPU."""

    def __init__(self, device: int = -1):
        self._L = lib()
        h = ctypes.c_void_p()
        rc = self._L.kzg_ctx_create(ctypes.byref(h), device)
        if rc != 0:
            raise KzgError(f"kzg_ctx_create failed with code {rc}: {self._L.kzg_getErrorName((1 << 64) - rc).decode()} "
                           "(a CUDA device is required; there is no CPU fallback)")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.kzg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ host memory in, host memory out
    def decompress_host(self, comp: np.ndarray, offsets: np.ndarray, sizes: np.ndarray, capacities: np.ndarray,
                        out: np.ndarray | None = None, out_offsets: np.ndarray | None = None):
        """Decodes frames given in host memory -> (out, out_offsets, status[u32], out_size[u64])."""
        n = int(len(offsets))
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
        capacities = np.ascontiguousarray(capacities, dtype=np.uint64)
        if out_offsets is None:
            out_offsets = np.zeros(n, np.uint64)
            if n:
                out_offsets[1:] = np.cumsum(capacities)[:-1]
        out_offsets = np.ascontiguousarray(out_offsets, dtype=np.uint64)
        if out is None:
            total = sum(int(c) for c in capacities)          # Python integers: a hostile capacity cannot wrap the sum
            if total >= 1 << 46:
                raise KzgError("capacities add up to more than any device holds")
            out = np.zeros(total, dtype=np.uint8)
        status = np.zeros(n, np.uint32)
        out_size = np.zeros(n, np.uint64)
        rc = self._L.kzg_decompress_batch_host(self._h, comp.ctypes.data, offsets.ctypes.data, sizes.ctypes.data,
                                               out.ctypes.data, out_offsets.ctypes.data, capacities.ctypes.data, n,
                                               status.ctypes.data, out_size.ctypes.data)
        if rc < 0:
            raise KzgError(f"kzg_decompress_batch_host failed: {rc}")
        return out, out_offsets, status, out_size

    # ------------------------------------------------------------------ device-resident batch
    def upload(self, comp: np.ndarray, offsets, sizes, capacities, device=None) -> DeviceBatch:
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        n = int(len(offsets))
        caps = np.ascontiguousarray(capacities, dtype=np.uint64)
        doff = np.zeros(n, np.uint64)
        if n:
            doff[1:] = np.cumsum(caps)[:-1]
        pad = 64

        def dev_u64(a):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(dev)

        src = torch.zeros(comp.size + pad, dtype=torch.uint8, device=dev)
        src[:comp.size] = torch.from_numpy(comp).to(dev)
        total_dst = sum(int(c) for c in caps)
        if total_dst >= 1 << 46:
            raise KzgError("capacities add up to more than any device holds")
        return DeviceBatch(src=src, src_off=dev_u64(offsets), src_size=dev_u64(sizes),
                           dst=torch.zeros(total_dst + pad, dtype=torch.uint8, device=dev), dst_off=dev_u64(doff),
                           dst_cap=dev_u64(caps), status=torch.zeros(n, dtype=torch.int32, device=dev),
                           out_size=torch.zeros(n, dtype=torch.int64, device=dev), n_frames=n,
                           total_src=int(np.asarray(sizes, dtype=np.uint64).sum()), total_dst=total_dst,
                           max_dst=int(caps.max()) if n else 0)

    def enqueue(self, b: DeviceBatch, stream=None) -> int:
        """Enqueues the decode of a device-resident batch on `stream` (torch stream or None = current)."""
        import torch
        s = (stream or torch.cuda.current_stream(b.src.device)).cuda_stream
        rc = self._L.kzg_decompress_batch_device(self._h, b.src.data_ptr(), b.src_off.data_ptr(), b.src_size.data_ptr(),
                                                 b.dst.data_ptr(), b.dst_off.data_ptr(), b.dst_cap.data_ptr(),
                                                 b.n_frames, b.total_src, b.total_dst, b.max_dst, b.status.data_ptr(),
                                                 b.out_size.data_ptr(), ctypes.c_void_p(s))
        if rc < 0:
            raise KzgError(f"kzg_decompress_batch_device failed: {rc}")
        return rc

    def finish(self) -> int:
        rc = self._L.kzg_batch_finish(self._h)
        if rc < 0:
            raise KzgError(f"kzg_batch_finish failed: {rc}")
        return rc

    def run(self, b: DeviceBatch, stream=None) -> int:
        """enqueue + finish, re-running while the optimistic workspace had to grow. Returns launches."""
        for _ in range(6):
            launches = self.enqueue(b, stream)
            if self.finish() == 0:
                return launches
        raise KzgError("workspace did not converge")

    def set_profiling(self, enable: bool = True) -> None:
        self._L.kzg_ctx_set_profiling(self._h, int(enable))

    def stage_ms(self, age: int = 0) -> dict:
        """Per-stage device time (ms) of the last finished batch (age 0) or of the batch enqueued `age` batches before
        it (the library keeps the events of the last 32 profiled batches); {} if profiling is off or nothing is kept."""
        ms = (ctypes.c_float * 16)()
        names = (ctypes.c_char_p * 16)()
        n = self._L.kzg_stage_ms_history(self._h, age, ms, names, 16)
        return {names[i].decode(): float(ms[i]) for i in range(max(n, 0))}

    def set_dictionary(self, dictionary: bytes | None) -> None:
        rc = self._L.kzg_ctx_set_dictionary(self._h, dictionary, len(dictionary) if dictionary else 0)
        if rc < 0:
            raise KzgError(f"kzg_ctx_set_dictionary failed: {rc}")

    def stats(self) -> BatchStats:
        st = BatchStats()
        self._L.kzg_last_batch_stats(self._h, ctypes.byref(st))
        return st


class BatchEncoder:
    """Batch frame writer on one GPU (include/kzg.h, kzg_compress_batch_*): every chunk of at most 128 KiB becomes
    one independent single-block zstd frame, the shape BatchDecoder decodes in bulk.  The step before the path
    (reference: ZstdWrapper.compressStream, ZstdWrapper.kt:26-46 -> Wrapper.cpp:75-121)."""

    def __init__(self, device: int = -1):
        self._L = lib()
        h = ctypes.c_void_p()
        rc = self._L.kzg_enc_create(ctypes.byref(h), device)
        if rc != 0:
            raise KzgError(f"kzg_enc_create failed with code {rc} (a CUDA device is required; there is no CPU fallback)")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.kzg_enc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bound(self, n: int) -> int:
        return int(self._L.kzg_compress_bound(n))

    def compress_chunks(self, data: np.ndarray, offsets: np.ndarray, sizes: np.ndarray, checksum: bool = False, level: int = 3):
        """Host memory in, host memory out -> (buffer, frame offsets, frame sizes, status).
        level < 3: greedy parse; level >= 3: one step of lazy matching (about 2 % smaller, match stage 25 % slower)."""
        n = int(len(offsets))
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
        caps = np.array([self.bound(int(s)) for s in sizes], dtype=np.uint64) if n else np.zeros(0, np.uint64)
        caps = (caps + 15) & ~np.uint64(15)
        doff = np.zeros(n, np.uint64)
        if n:
            doff[1:] = np.cumsum(caps)[:-1]
        out = np.zeros(int(caps.sum()) + 64, np.uint8)
        status = np.zeros(n, np.uint32)
        out_size = np.zeros(n, np.uint64)
        rc = self._L.kzg_compress_batch_host(self._h, data.ctypes.data if data.size else None, offsets.ctypes.data,
                                             sizes.ctypes.data, out.ctypes.data, doff.ctypes.data, caps.ctypes.data, n,
                                             (1 if checksum else 0) | (2 if level >= 3 else 0), status.ctypes.data, out_size.ctypes.data)
        if rc < 0:
            raise KzgError(f"kzg_compress_batch_host failed: {rc}")
        return out, doff, out_size, status

    def compress(self, data, chunk_size: int = 65536, checksum: bool = False, level: int = 3, seek_table: bool = False) -> bytes:
        """Cuts `data` into chunks of chunk_size and returns the concatenated frames (a valid zstd stream: any
        decoder yields `data` again)."""
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        if not 0 < chunk_size <= 131072:
            raise ValueError("chunk_size must be in (0, 131072]")
        n = max(1, (a.size + chunk_size - 1) // chunk_size)
        offsets = np.arange(n, dtype=np.uint64) * np.uint64(chunk_size)
        sizes = np.minimum(np.uint64(chunk_size), np.uint64(a.size) - offsets).astype(np.uint64) if a.size else np.zeros(1, np.uint64)
        out, doff, osz, status = self.compress_chunks(a, offsets, sizes, checksum, level)
        if status.any():
            raise KzgError(f"compress failed: status {status[status != 0][0]}")
        frames = b"".join(out[int(o):int(o) + int(s)].tobytes() for o, s in zip(doff, osz))
        if not seek_table:
            return frames
        # the seekable format's index: a skippable frame that lists every frame's compressed and content size
        cap = int(self._L.kzg_seek_table_bound(n))
        tab = np.zeros(cap, np.uint8)
        osz64 = np.ascontiguousarray(osz, dtype=np.uint64)
        wrote = self._L.kzg_write_seek_table(tab.ctypes.data, cap, osz64.ctypes.data, sizes.ctypes.data, n)
        if self._L.kzg_isError(wrote):
            raise KzgError("kzg_write_seek_table failed")
        return frames + tab[:wrote].tobytes()

    def set_profiling(self, enable: bool = True) -> None:
        self._L.kzg_enc_set_profiling(self._h, int(enable))

    def stage_ms(self) -> dict:
        ms = (ctypes.c_float * 4)()
        names = (ctypes.c_char_p * 4)()
        n = self._L.kzg_enc_last_stage_ms(self._h, ms, names, 4)
        return {names[i].decode(): float(ms[i]) for i in range(max(n, 0))}
