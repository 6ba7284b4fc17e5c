"""Differential stress test on a GPU (test infrastructure, not part of the default suite; run from the repository root):
   python tests/stress/gpu_fuzz.py [n_frames] [seed]
Random inputs (all corpus kinds, run-like data, random sizes 0..400 KB, levels -5..19, checksum on/off, pledged
and streamed headers) are compressed by the real libzstd and decoded in batches through the C ABI; then single-bit
corruptions are checked for the same accept/reject decision (and bytes) as the oracle."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import corpus, oracle
from oracle.libzstd_ref import load
from kompressor_b200.batch import BatchDecoder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
z = load("1.5.7"); oracle.build()
dec = BatchDecoder()
kinds = ["logs", "json", "literal", "random", "mixed"]
raws, frames = [], []
t0 = time.time()
for i in range(n):
    r = rng.random()
    size = int(rng.integers(0, 2000)) if r < 0.2 else int(rng.integers(0, 70000)) if r < 0.7 else int(rng.integers(0, 400000))
    k = rng.integers(0, 8)
    if k < 5: raw = bytes(corpus.generate(kinds[k], int(rng.integers(1, 1 << 30)), size))
    elif k == 5: raw = bytes(rng.integers(0, int(rng.integers(1, 5)), size, dtype=np.uint8))          # tiny alphabet: runs
    elif k == 6: raw = (bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8)) * (size // 8 + 1))[:size]   # short period
    else:
        a = bytes(corpus.generate("logs", int(rng.integers(1, 1 << 30)), size // 2)); raw = (a + a[::-1] + a)[:size]
    lvl = int(rng.choice([-5, 1, 2, 3, 3, 3, 5, 7, 9, 12, 16, 19]))
    frames.append(z.compress(raw, lvl, checksum=bool(rng.integers(0, 2)), pledged=bool(rng.integers(0, 2))))
    raws.append(raw)
print(f"built {n} frames in {time.time()-t0:.1f}s, {sum(map(len, raws))/1e6:.1f} MB raw", flush=True)

def run(fr, caps):
    blob = b"".join(fr)
    comp = np.frombuffer(blob, dtype=np.uint8).copy() if blob else np.zeros(1, np.uint8)
    sizes = np.array([len(f) for f in fr], dtype=np.uint64)
    offs = np.zeros(len(fr), np.uint64); offs[1:] = np.cumsum(sizes)[:-1]
    out, doff, status, osz = dec.decompress_host(comp, offs, sizes, np.array(caps, dtype=np.uint64))
    return [bytes(out[int(doff[i]):int(doff[i] + osz[i])]) for i in range(len(fr))], status

bad = 0
for lo in range(0, n, 200):
    outs, status = run(frames[lo:lo + 200], [len(r) for r in raws[lo:lo + 200]])
    for i, (o, s) in enumerate(zip(outs, status)):
        if s != 0 or o != raws[lo + i]:
            bad += 1; print("MISMATCH frame", lo + i, "status", s, "len", len(raws[lo + i]), flush=True)
print("round trips:", "OK" if bad == 0 else f"{bad} FAILED", flush=True)
# corruptions
nb = 0; ncorr = min(n, 400) * 2
cf = []
for j in range(ncorr):
    f = bytearray(frames[j % len(frames)])
    if len(f) < 8: cf.append(bytes(f)); continue
    for _ in range(int(rng.integers(1, 3))): f[int(rng.integers(4, len(f)))] ^= 1 << int(rng.integers(0, 8))
    cf.append(bytes(f))
outs, status = run(cf, [len(raws[j % len(frames)]) + 64 for j in range(ncorr)])
for j, (f, o, s) in enumerate(zip(cf, outs, status)):
    ro, re_ = oracle.decompress(f, len(raws[j % len(frames)]) + 64)
    if (re_ == 0) != (s == 0) or (re_ == 0 and o != ro):
        nb += 1; print("CORRUPT-PARITY mismatch", j, "oracle", re_, "gpu", s, flush=True)
print("corruptions:", "OK" if nb == 0 else f"{nb} FAILED", f"({int((status == 0).sum())} of {ncorr} still decode)", flush=True)
sys.exit(0 if bad == 0 and nb == 0 else 1)
