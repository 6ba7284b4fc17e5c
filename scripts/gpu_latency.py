import sys, time, ctypes
sys.path.insert(0, ".")
import numpy as np, corpus
from kompressor_b200._lib import lib
from kompressor_b200.batch import BatchDecoder
L = lib()
c = corpus.build("logs", 512, 65536, 3, False)
comp = np.ascontiguousarray(c.comp)
dec = BatchDecoder()
L.kzg_ctx_set_profiling(dec._h, 1)
names = (ctypes.c_char_p * 16)(); ms = (ctypes.c_float * 16)()
for n in (512, 64, 8, 1, 1, 8, 31, 32, 33, 64):
    offs = np.ascontiguousarray(c.offsets[:n], dtype=np.uint64); sizes = np.ascontiguousarray(c.sizes[:n], dtype=np.uint64)
    caps = np.full(n, 65536, np.uint64); dst = np.empty(n * 65536 + 64, np.uint8); doff = np.arange(n, dtype=np.uint64) * 65536
    for _ in range(20): dec.decompress_host(comp, offs, sizes, caps, out=dst, out_offsets=doff)
    ts = []
    for _ in range(10):
        t = time.perf_counter(); dec.decompress_host(comp, offs, sizes, caps, out=dst, out_offsets=doff); ts.append(time.perf_counter() - t)
    k = L.kzg_last_stage_ms(dec._h, ms, names, 16)
    print(f"batch {n:4d}: {min(ts)*1e3:7.3f} ms | " + " ".join(f"{names[i].decode()}={ms[i]:.3f}" for i in range(k)), flush=True)
