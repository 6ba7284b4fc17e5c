"""Quick GPU parity check used during development: python scripts/gpu_check.py [n_frames]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import corpus
from kompressor_b200.batch import BatchDecoder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dec = BatchDecoder()
ok_all = True
for kind, lvl, cs in [("random", 3, False), ("logs", 3, False), ("json", 3, True), ("literal", 1, False)]:
    c = corpus.build(kind, n, 65536, lvl, cs)
    caps = np.full(n, 65536, np.uint64)
    out, doff, status, osz = dec.decompress_host(c.comp, c.offsets, c.sizes, caps)
    ok = bool((status == 0).all()) and bool((osz == 65536).all()) and np.array_equal(out, c.raw)
    nbad = int((out.reshape(n, 65536) != c.raw.reshape(n, 65536)).any(axis=1).sum())
    print(kind, "OK" if ok else "MISMATCH", "status", np.unique(status), "bad frames", nbad, flush=True)
    ok_all &= ok
    b = dec.upload(c.comp, c.offsets, c.sizes, caps)
    dec.run(b)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        ev0.record(); dec.enqueue(b); ev1.record(); dec.finish(); ts.append(ev0.elapsed_time(ev1))
    st = dec.stats()
    ok2 = bool((b.status == 0).all()) and np.array_equal(b.dst[:n * 65536].cpu().numpy(), c.raw)
    print(f"  device: {min(ts):.3f} ms  -> {n*65536/min(ts)/1e6:.1f} GB/s out; blocks {st.n_blocks} cblocks {st.n_cblocks} ok={ok2}", flush=True)
    ok_all &= ok2
print("ALL OK" if ok_all else "FAILED")
sys.exit(0 if ok_all else 1)
