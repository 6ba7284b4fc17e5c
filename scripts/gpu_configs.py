"""Timing of BASELINE.json's other configs on one GPU (development):
   python scripts/gpu_configs.py [scale] [kind ...]   (scale 1.0 = 1 GiB per config, except level 19; kinds: all; mixed:64 and mixed:148 = configs[4] with more frames in the batch, only on request)"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import corpus
from kompressor_b200.batch import BatchDecoder

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
only = set(sys.argv[2:])
dec = BatchDecoder()
dec.set_profiling(True)
cases = [("logs", 65536, 3, False, False, int(16384 * scale)), ("json", 65536, 3, True, False, int(16384 * scale)),
         ("literal", 65536, 1, False, False, int(16384 * scale)), ("random", 131075, 3, False, False, 256),
         ("mixed", 4 << 20, 19, False, True, 16), ("mixed:64", 4 << 20, 19, False, True, 64), ("mixed:148", 4 << 20, 19, False, True, 148)]
for kind, fs, lvl, cs, streamed, n in cases:
    if (only and kind not in only) or (not only and ":" in kind):      # the larger level-19 batches only on request
        continue
    kind = kind.split(":")[0]
    t = time.time()
    c = corpus.build(kind, n, fs, lvl, cs, streamed=streamed)
    caps = np.full(n, fs, np.uint64)
    b = dec.upload(c.comp, c.offsets, c.sizes, caps)
    dec.run(b)
    ok = bool((b.status == 0).all()) and np.array_equal(b.dst[:n * fs].cpu().numpy(), c.raw)
    ts = []
    for _ in range(5):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); dec.enqueue(b); ev1.record(); dec.finish(); ts.append(ev0.elapsed_time(ev1))
    best = min(ts)
    st = {k: round(v, 3) for k, v in dec.stage_ms().items()}
    print(f"{kind:8s} fs={fs} lvl={lvl} cs={int(cs)} n={n} ratio={c.ratio:.2f} build={time.time()-t:.0f}s parity={ok} "
          f"best={best:.3f} ms -> {n*fs/best/1e6:.1f} GB/s out  stages={st}", flush=True)
