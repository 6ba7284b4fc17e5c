"""Differential stress test of the second execute design on the CPU (test infrastructure; run from the repository root):
   python tests/stress/emu_fuzz.py [n_frames] [seed] [tile]
The kernel source kzg_exec2.cuh runs on the lock-step warp emulator (tests/hostsim/simt_emu.h) behind the host-compiled
stage code.  Inputs as in gpu_fuzz.py (all corpus kinds, run-like and short-period data, sizes 0..400 KB, levels -5..19,
checksums, pledged and streamed headers), compressed by the real libzstd; then bit corruptions are checked against the
oracle's accept / reject decision and bytes."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import corpus, oracle, hostsim
from oracle.libzstd_ref import load

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
rng = np.random.default_rng(seed)
z = load("1.5.7"); oracle.build()
kinds = ["logs", "json", "literal", "random", "mixed"]
raws, frames = [], []
t0 = time.time()
for i in range(n):
    r = rng.random()
    size = int(rng.integers(0, 2000)) if r < 0.3 else int(rng.integers(0, 70000)) if r < 0.85 else int(rng.integers(0, 400000))
    k = rng.integers(0, 9)
    if k < 5: raw = bytes(corpus.generate(kinds[k], int(rng.integers(1, 1 << 30)), size))
    elif k == 5: raw = bytes(rng.integers(0, int(rng.integers(1, 5)), size, dtype=np.uint8))          # tiny alphabet: runs
    elif k == 6: raw = (bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8)) * (size // 8 + 1))[:size]   # short period
    elif k == 7:
        a = bytes(corpus.generate("logs", int(rng.integers(1, 1 << 30)), size // 2)); raw = (a + a[::-1] + a)[:size]   # long far matches
    else:
        parts = []                                                                                      # long literal runs between matches
        while sum(map(len, parts)) < size:
            parts.append(bytes(rng.integers(0, 256, int(rng.integers(1, 3000)), dtype=np.uint8)))
            parts.append(parts[int(rng.integers(0, len(parts)))][:int(rng.integers(3, 400))])
        raw = b"".join(parts)[:size]
    lvl = int(rng.choice([-5, 1, 2, 3, 3, 3, 5, 7, 9, 12, 16, 19]))
    frames.append(z.compress(raw, lvl, checksum=bool(rng.integers(0, 2)), pledged=bool(rng.integers(0, 2))))
    raws.append(raw)
print(f"built {n} frames in {time.time()-t0:.1f}s, {sum(map(len, raws))/1e6:.1f} MB raw", flush=True)
bad = 0
for mode in (1, 2):
    t0 = time.time()
    outs, status, _ = hostsim.run(frames, [len(r) + (i % 5) for i, r in enumerate(raws)], exec_mode=mode, tile=tile)
    for i, (o, s) in enumerate(zip(outs, status)):
        if s != 0 or o != raws[i]:
            bad += 1; print("MISMATCH frame", i, "mode", mode, "status", s, "len", len(raws[i]), flush=True)
    print(f"mode {mode}: {time.time()-t0:.1f}s, deferred so far {hostsim._lib.kzs_exec_deferred()}", flush=True)
print("round trips:", "OK" if bad == 0 else f"{bad} FAILED", flush=True)
nb = 0; ncorr = min(n, 300) * 2
cf = []
for j in range(ncorr):
    f = bytearray(frames[j % len(frames)])
    if len(f) < 8: cf.append(bytes(f)); continue
    for _ in range(int(rng.integers(1, 3))): f[int(rng.integers(4, len(f)))] ^= 1 << int(rng.integers(0, 8))
    cf.append(bytes(f))
outs, status, _ = hostsim.run(cf, [len(raws[j % len(frames)]) + 64 for j in range(ncorr)], exec_mode=1, tile=tile)
for j, (f, o, s) in enumerate(zip(cf, outs, status)):
    ro, re_ = oracle.decompress(f, len(raws[j % len(frames)]) + 64)
    if (re_ == 0) != (s == 0) or (re_ == 0 and o != ro):
        nb += 1; print("CORRUPT-PARITY mismatch", j, "oracle", re_, "emu", s, flush=True)
print("corruptions:", "OK" if nb == 0 else f"{nb} FAILED", f"({int((status == 0).sum())} of {ncorr} still decode)", flush=True)
sys.exit(0 if bad == 0 and nb == 0 else 1)
