"""Batch frame writer: ratio, device time per stage, round trip (development / DESIGN.md numbers).
python scripts/gpu_compress_bench.py [frames] [kind]"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

import corpus
from kompressor_b200._lib import lib
from kompressor_b200.batch import BatchDecoder, BatchEncoder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kinds = sys.argv[2:] or ["logs", "json", "literal"]
L = lib()
for kind in kinds:
    c = corpus.build(kind, n, 65536, 3 if kind != "literal" else 1, False)
    raw = np.ascontiguousarray(c.raw)
    enc = BatchEncoder()
    enc.set_profiling(True)
    dev = torch.device("cuda", 0)
    d_src = torch.zeros(raw.size + 64, dtype=torch.uint8, device=dev)
    d_src[: raw.size] = torch.from_numpy(raw).to(dev)
    cap = (enc.bound(65536) + 15) & ~15
    offs = torch.arange(n, dtype=torch.int64, device=dev) * 65536
    sizes = torch.full((n,), 65536, dtype=torch.int64, device=dev)
    doff = torch.arange(n, dtype=torch.int64, device=dev) * cap
    dcap = torch.full((n,), cap, dtype=torch.int64, device=dev)
    d_dst = torch.zeros(n * cap + 64, dtype=torch.uint8, device=dev)
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    osz = torch.zeros(n, dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    import ctypes

    def run(flags):
        rc = L.kzg_compress_batch_device(enc._h, d_src.data_ptr(), offs.data_ptr(), sizes.data_ptr(), d_dst.data_ptr(),
                                         doff.data_ptr(), dcap.data_ptr(), n, 65536, flags, status.data_ptr(), osz.data_ptr(),
                                         ctypes.c_void_p(s))
        assert rc > 0, rc
        assert L.kzg_enc_finish(enc._h, ctypes.c_void_p(s)) == 0
    for flags, name in ((0, "greedy (level < 3)"), (2, "lazy (level >= 3)")):
        for _ in range(2):
            run(flags)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(5):
            e0.record(); run(flags); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        assert int(status.abs().sum()) == 0
        comp_bytes = int(osz.sum())
        ms = min(ts)
        print(f"{kind}, {name}: {n} chunks of 64 KiB: ratio {raw.size / comp_bytes:.3f} (libzstd level {3 if kind != 'literal' else 1}: "
              f"{raw.size / c.comp.size:.3f}); {ms:.2f} ms = {raw.size / ms / 1e6:.1f} GB/s in; stages {enc.stage_ms()}", flush=True)
    # round trip through the GPU decoder (device resident)
    dec = BatchDecoder()
    b = dec.upload(np.zeros(1, np.uint8), [0], [1], [1])      # placeholder to get the types
    from kompressor_b200.batch import DeviceBatch
    out = torch.zeros(n * 65536 + 64, dtype=torch.uint8, device=dev)
    db = DeviceBatch(src=d_dst, src_off=doff, src_size=osz, dst=out, dst_off=offs, dst_cap=sizes,
                     status=torch.zeros(n, dtype=torch.int32, device=dev), out_size=torch.zeros(n, dtype=torch.int64, device=dev),
                     n_frames=n, total_src=comp_bytes, total_dst=n * 65536, max_dst=65536)
    dec.run(db)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); dec.run(db); torch.cuda.synchronize(); t1 = time.perf_counter()
    ok = bool(torch.equal(out[: raw.size].cpu(), torch.from_numpy(raw))) and int(db.status.abs().sum()) == 0
    print(f"   GPU decode of the GPU frames: parity {ok}, {(t1 - t0) * 1e3:.2f} ms", flush=True)

# ---- host memory in, host memory out (kzg_compress_batch_host, pinned buffers): H2D, compress and D2H pipelined
kind = kinds[0]
c = corpus.build(kind, n, 65536, 3, False)
enc = BatchEncoder()
raw_p = torch.from_numpy(np.ascontiguousarray(c.raw)).pin_memory().numpy()
cap = (enc.bound(65536) + 15) & ~15
h_off = np.arange(n, dtype=np.uint64) * 65536
h_size = np.full(n, 65536, np.uint64)
h_doff = np.arange(n, dtype=np.uint64) * cap
h_cap = np.full(n, cap, np.uint64)
h_dst = torch.empty(n * cap + 64, dtype=torch.uint8).pin_memory().numpy()
h_status = np.zeros(n, np.uint32)
h_osz = np.zeros(n, np.uint64)
ts = []
for _ in range(6):
    t0 = time.perf_counter()
    rc = L.kzg_compress_batch_host(enc._h, raw_p.ctypes.data, h_off.ctypes.data, h_size.ctypes.data, h_dst.ctypes.data,
                                   h_doff.ctypes.data, h_cap.ctypes.data, n, 2, h_status.ctypes.data, h_osz.ctypes.data)
    ts.append(time.perf_counter() - t0)
    assert rc == 0 and not h_status.any()
print(f"host API ({kind}, lazy, {n} chunks, pinned): best {min(ts[1:]) * 1e3:.2f} ms = {raw_p.size / min(ts[1:]) / 1e9:.1f} GB/s in", flush=True)
