"""Development timing: python scripts/gpu_time.py kind n_frames [level] [checksum] [iters]"""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import corpus
from kompressor_b200.batch import BatchDecoder

kind = sys.argv[1] if len(sys.argv) > 1 else "logs"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
lvl = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cs = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 10
t = time.time()
c = corpus.build(kind, n, 65536, lvl, cs)
print(f"corpus {kind}: {n} frames ratio {c.ratio:.2f} built in {time.time()-t:.1f}s on {os.cpu_count()} cpus", flush=True)
dec = BatchDecoder()
caps = np.full(n, 65536, np.uint64)
b = dec.upload(c.comp, c.offsets, c.sizes, caps)
dec.run(b)
ok = bool((b.status == 0).all()) and np.array_equal(b.dst[:n * 65536].cpu().numpy(), c.raw)
print("parity", ok, flush=True)
ts = []
for _ in range(iters):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); dec.enqueue(b); ev1.record(); dec.finish(); ts.append(ev0.elapsed_time(ev1))
st = dec.stats()
best = min(ts)
print(f"device best {best:.3f} ms median {sorted(ts)[len(ts)//2]:.3f} ms -> {n*65536/best/1e6:.1f} GB/s out, "
      f"{(n*65536+c.comp_bytes)/best/1e6:.1f} GB/s algorithmic = {(n*65536+c.comp_bytes)/best/1e6/6533.2*100:.2f}% of 6533 GB/s; "
      f"ws {st.workspace_bytes/2**20:.0f} MiB seqs {st.seq_records} lit {st.lit_bytes}", flush=True)
sys.exit(0 if ok else 1)
