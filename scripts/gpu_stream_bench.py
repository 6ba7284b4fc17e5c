"""Throughput of the streaming shim (layer B) in the reference's call patterns (development / DESIGN.md numbers):

  transform  SliceTransform.transform(ByteArray) (SliceTransform.kt:33-45): whole input offered in the first call,
             output slices of max(8192, n/10)
  pipe       SliceTransformRawSource.readAtMostTo (SliceTransformRawSource.kt:32-55): 8 KiB in, 8 KiB out
  pipe xT    T threads, one handle each (one HTTP body each), same pattern
  batch n    kzg_decompress_batch_host latency for n frames, with the per-stage times

python scripts/gpu_stream_bench.py [frames_for_transform] [frames_per_pipe_handle] [--native]
--native: the same patterns driven by scripts/stream_bench.cpp (native threads, no interpreter in the loops).
"""
import ctypes
import sys
import threading
import time

sys.path.insert(0, ".")
import numpy as np

import corpus
from kompressor_b200._lib import InBuffer, OutBuffer, lib
from kompressor_b200.batch import BatchDecoder

L = lib()
NATIVE = "--native" in sys.argv
_args = [a for a in sys.argv[1:] if not a.startswith("--")]
NT = int(_args[0]) if len(_args) > 0 else 2048
NP = int(_args[1]) if len(_args) > 1 else 64


def native(c):
    import os
    import subprocess
    os.makedirs("gpurun_out", exist_ok=True)
    path = "/tmp/kzg_stream_corpus.bin"
    n = len(c.offsets)
    off = np.zeros(n + 1, np.uint64)
    off[:n] = c.offsets
    off[n] = c.comp.size
    with open(path, "wb") as f:
        f.write(np.array([n, c.comp.size, c.raw.size], np.uint64).tobytes())
        f.write(off.tobytes())
        f.write(np.ascontiguousarray(c.comp).tobytes())
        f.write(np.ascontiguousarray(c.raw).tobytes())
    exe = "/tmp/kzg_stream_bench"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "scripts/stream_bench.cpp", "-Iinclude", "-Lkompressor_b200",
                           "-lkzg", "-Wl,-rpath," + os.path.abspath("kompressor_b200"), "-o", exe])
    subprocess.check_call([exe, path, str(NP)])


def run_stream(comp: np.ndarray, raw_len: int, in_slice: int, out_slice: int, out: np.ndarray) -> int:
    """Drives one handle the way the Kotlin loops do. in_slice = 0: everything in the first call."""
    d = L.kzg_createDCtx()
    assert d
    n = comp.size
    ip = 0
    op = 0
    base_in = comp.ctypes.data
    base_out = out.ctypes.data
    inb = InBuffer()
    outb = OutBuffer()
    while True:
        take = n - ip if in_slice == 0 else min(in_slice, n - ip)
        inb.src, inb.size, inb.pos = base_in + ip, take, 0
        while True:
            room = min(out_slice, out.size - op)
            outb.dst, outb.size, outb.pos = base_out + op, room, 0
            r = L.kzg_decompressStream(d, ctypes.byref(outb), ctypes.byref(inb))
            assert not L.kzg_isError(r), L.kzg_getErrorName(r)
            op += outb.pos
            # SliceTransform.kt:39-45: while (input.remainingRead != 0 || output.insufficient), insufficient = isFull && result != 0
            if not (inb.pos < inb.size or (outb.pos == room and r != 0)):
                break
        ip += take
        if ip >= n:
            break
    L.kzg_freeDCtx(d)
    return op


def timed(label, fn, nbytes, reps=3):
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    print(f"{label:34s} {best * 1e3:9.2f} ms  {nbytes / best / 1e6:10.1f} MB/s", flush=True)


def main():
    c = corpus.build("logs", NT, 65536, 3, False)
    if NATIVE:
        return native(c)
    comp = np.ascontiguousarray(c.comp)
    out = np.empty(c.raw.size + 64, np.uint8)
    run_stream(comp[: int(c.offsets[8])], 8 * 65536, 0, 65536, out)       # warm-up (context, workspace)

    def tr():
        got = run_stream(comp, c.raw.size, 0, max(8192, comp.size // 10), out)
        assert got == c.raw.size
    timed(f"transform {NT} frames", tr, c.raw.size)
    assert np.array_equal(out[: c.raw.size], c.raw)

    small = np.ascontiguousarray(comp[: int(c.offsets[NP])])
    small_raw = NP * 65536
    out[:] = 0

    def pipe1():
        got = run_stream(small, small_raw, 8192, 8192, out)
        assert got == small_raw
    timed(f"pipe 8K/8K, 1 handle, {NP} frames", pipe1, small_raw)
    assert np.array_equal(out[:small_raw], c.raw[:small_raw])

    for T in (4, 16, 64):
        outs = [np.empty(small_raw + 64, np.uint8) for _ in range(T)]

        def pipeT():
            th = [threading.Thread(target=run_stream, args=(small, small_raw, 8192, 8192, outs[i])) for i in range(T)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        timed(f"pipe 8K/8K, {T} handles x {NP} frames", pipeT, small_raw * T, reps=2)
        for o in outs:
            assert np.array_equal(o[:small_raw], c.raw[:small_raw])

    dec = BatchDecoder()
    L.kzg_ctx_set_profiling(dec._h, 1)
    names = (ctypes.c_char_p * 16)()
    ms = (ctypes.c_float * 16)()
    for n in (1, 8, 64, 512):
        offs = np.ascontiguousarray(c.offsets[:n], dtype=np.uint64)
        sizes = np.ascontiguousarray(c.sizes[:n], dtype=np.uint64)
        caps = np.full(n, 65536, np.uint64)
        dst = np.empty(n * 65536 + 64, np.uint8)
        doff = np.arange(n, dtype=np.uint64) * 65536
        for _ in range(2):
            dec.decompress_host(comp, offs, sizes, caps, out=dst, out_offsets=doff)
        ts = []
        for _ in range(5):
            t = time.perf_counter()
            dec.decompress_host(comp, offs, sizes, caps, out=dst, out_offsets=doff)
            ts.append(time.perf_counter() - t)
        k = L.kzg_last_stage_ms(dec._h, ms, names, 16)
        st = " ".join(f"{names[i].decode()}={ms[i]:.3f}" for i in range(k))
        print(f"batch {n:4d} frames: {min(ts) * 1e3:7.3f} ms host wall | {st}", flush=True)


if __name__ == "__main__":
    main()
