"""Development aid (GPU): one pass of BASELINE configs[4] (16 frames of 4 MiB, level 19, streamed) for profiling."""
import sys; sys.path.insert(0, ".")
import numpy as np, torch, corpus
from kompressor_b200.batch import BatchDecoder
dec = BatchDecoder(); dec.set_profiling(True)
n, fs = 16, 4 << 20
c = corpus.build("mixed", n, fs, 19, False, streamed=True)
b = dec.upload(c.comp, c.offsets, c.sizes, np.full(n, fs, np.uint64))
dec.run(b)
print("parity", bool((b.status == 0).all()) and np.array_equal(b.dst[:n * fs].cpu().numpy(), c.raw))
dec.enqueue(b); dec.finish(); print(dec.stage_ms())
