"""e2e timing of kzg_decompress_batch_host (development): python scripts/gpu_e2e.py [frames]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch, corpus
from kompressor_b200.batch import BatchDecoder
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
c = corpus.build("logs", n, 65536, 3, False)
dec = BatchDecoder()
caps = np.full(n, 65536, np.uint64)
src = torch.from_numpy(c.comp).pin_memory().numpy(); dst = torch.empty(n * 65536, dtype=torch.uint8).pin_memory().numpy()
doff = np.arange(n, dtype=np.uint64) * 65536
offs = np.ascontiguousarray(c.offsets, dtype=np.uint64); sizes = np.ascontiguousarray(c.sizes, dtype=np.uint64)
for _ in range(2): dec.decompress_host(src, offs, sizes, caps, out=dst, out_offsets=doff)
ts = []
for _ in range(6):
    t = time.perf_counter(); dec.decompress_host(src, offs, sizes, caps, out=dst, out_offsets=doff); ts.append(time.perf_counter() - t)
ok = np.array_equal(dst, c.raw)
print(f"e2e best {min(ts)*1e3:.2f} ms median {sorted(ts)[3]*1e3:.2f} ms -> {n*65536/min(ts)/1e9:.1f} GB/s parity {ok}")
