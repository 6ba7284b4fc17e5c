"""Stress test of the batch frame writer on a GPU (test infrastructure, not part of the default suite; run from the repository root):
   python tests/stress/gpu_fuzz_compress.py [n_inputs] [seed]
Random inputs (all corpus kinds, tiny alphabets, short periods, mirrored text; sizes 0..400 KB; chunk sizes 1..128 KiB;
checksum on/off) are compressed on the GPU; the real libzstd must decode every result to the input, and so must the
GPU decoder."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import corpus
from oracle.libzstd_ref import load
from kompressor_b200.batch import BatchEncoder
from kompressor_b200.mirror import core, zstd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
z = load("1.5.7")
enc = BatchEncoder()
kinds = ["logs", "json", "literal", "random", "mixed"]
bad = 0
total_in = total_out = 0
t0 = time.time()
for i in range(n):
    r = rng.random()
    size = int(rng.integers(0, 2000)) if r < 0.2 else int(rng.integers(0, 70000)) if r < 0.6 else int(rng.integers(0, 400000))
    k = rng.integers(0, 9)
    if k < 5: raw = bytes(corpus.generate(kinds[k], int(rng.integers(1, 1 << 30)), size))
    elif k == 5: raw = bytes(rng.integers(0, int(rng.integers(1, 5)), size, dtype=np.uint8))
    elif k == 6: raw = (bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8)) * (size // 8 + 1))[:size]
    elif k == 7: raw = bytes(rng.integers(32, int(rng.integers(34, 127)), size, dtype=np.uint8))       # printable, few symbols
    else:
        a = bytes(corpus.generate("logs", int(rng.integers(1, 1 << 30)), size // 2)); raw = (a + a[::-1] + a)[:size]
    chunk = int(rng.choice([65536, 65536, 131072, 32768, 4096, 1000, 100000, int(rng.integers(1, 131073))]))
    cs = bool(rng.integers(0, 2))
    comp = enc.compress(raw, chunk_size=chunk, checksum=cs)
    out, err = z.decompress_oneshot(comp, len(raw) + 64)
    ok = err == 0 and out == raw
    if ok and i % 4 == 0:
        ok = core.transform(zstd.ZstdDecompressor(), comp) == raw
    if not ok:
        bad += 1
        print("MISMATCH input", i, "kind", int(k), "size", size, "chunk", chunk, "checksum", cs, "err", err, flush=True)
    total_in += len(raw); total_out += len(comp)
print(f"{n} inputs, {total_in / 1e6:.1f} MB -> {total_out / 1e6:.1f} MB in {time.time() - t0:.1f}s:", "OK" if bad == 0 else f"{bad} FAILED", flush=True)
sys.exit(1 if bad else 0)
