"""The third execute design (kompressor_b200/csrc/kzg_exec3.cuh: a thread per frame, word-interleaved ring windows, per-lane
entry FIFO with cp.async staging, transposed warp flush) run on the lock-step warp emulator (tests/hostsim/simt_emu.h): the kernel SOURCE, compiled
for the host, 32 coroutines per warp, asynchronous copies completed as late as their waits allow, lanes scheduled in
forward and in reverse order.  Host logic testing only -- the GPU parity tests are in test_gpu_*.py.
What it stands in for: ZSTD_execSequence under ZSTD_decompressStream
(kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:178)."""
import hashlib

import numpy as np
import pytest

from helpers import check_output, frames_of, gbytes

hostsim = pytest.importorskip("hostsim")


@pytest.mark.parametrize("mode", [3, 4, 5])
def test_golden_vectors(manifest, mode):
    vs = [v for v in manifest["vectors"] if "dictionary_b64" not in v and v["name"] not in ("skippable_then_hello", "two_frames")]
    outs, status, _ = hostsim.run([gbytes(v["name"] + ".zst") for v in vs], [v["size"] + 32 for v in vs], exec_mode=mode, tile=1024)
    for v, o, s in zip(vs, outs, status):
        assert s == 0, v["name"]
        check_output(v, o)


@pytest.mark.parametrize("tile", [512, 1024, 2048])
@pytest.mark.parametrize("kind,level,checksum", [("logs", 3, False), ("json", 3, True), ("literal", 1, False)])
def test_corpora(kind, level, checksum, tile):
    import corpus
    c = corpus.build(kind, 6, 65536, level, checksum)
    for mode in (3, 4, 5, 6):
        # ragged capacities: every frame starts at a different alignment of the output buffer
        outs, status, sizes = hostsim.run(frames_of(c), [65536 + (i % 7) for i in range(c.n_frames)], exec_mode=mode, tile=tile)
        assert (status == 0).all() and (sizes == 65536).all()
        assert b"".join(outs) == c.raw.tobytes()


def test_multiblock_long_offsets(libzstd):
    """BASELINE configs[4] in small: level 19, streamed header (window descriptor, no content size), logs + JSON."""
    import corpus
    raw = bytes(corpus.generate("mixed", 77, 600000))
    frames = [libzstd.compress(raw, lvl, checksum=True, pledged=p) for lvl, p in ((19, False), (3, True))]
    for tile in (512, 1024):
        outs, status, _ = hostsim.run(frames, [len(raw) + 5] * len(frames), exec_mode=3, tile=tile)
        assert (status == 0).all()
        for o in outs:
            assert hashlib.sha256(o).digest() == hashlib.sha256(raw).digest()


def test_runs_periods_and_long_copies(libzstd):
    """Overlapping matches of every short period, long matches from far behind, long literal runs, tiny inputs."""
    import corpus
    rng = np.random.default_rng(11)
    raws = [b"", b"a", b"ab" * 2, bytes(40000), b"abc" * 20000]
    raws += [(bytes(rng.integers(0, 256, p, dtype=np.uint8)) * (30000 // p + 1))[:30000 + p] for p in (1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 33, 100)]
    a = bytes(corpus.generate("logs", 9, 50000))
    raws.append(a + a[::-1] + a)                                                     # matches of tens of kilobytes from 100 KB behind
    raws.append(bytes(rng.integers(0, 3, 70000, dtype=np.uint8)))                    # tiny alphabet
    raws += [bytes(corpus.generate("logs", 20 + i, n)) for i, n in enumerate([1, 2, 3, 5, 17, 100, 1000, 4097, 65535, 65537])]
    frames = [libzstd.compress(r, lvl) for r in raws for lvl in (3,)] + [libzstd.compress(r, 19) for r in raws[-4:]]
    exp = raws + raws[-4:]
    for mode in (3, 4, 5, 6):
        outs, status, _ = hostsim.run(frames, [len(r) + 3 for r in exp], exec_mode=mode, tile=1024)
        assert (status == 0).all()
        assert outs == exp


def test_corrupt_sequences_are_rejected_like_the_oracle(oracle_mod, libzstd):
    import corpus
    rng = np.random.default_rng(5)
    raw = bytes(corpus.generate("json", 3, 20000))
    frame = libzstd.compress(raw, 3)
    flips = []
    for _ in range(120):
        f = bytearray(frame)
        f[int(rng.integers(8, len(f)))] ^= 1 << int(rng.integers(0, 8))
        flips.append(bytes(f))
    outs, status, _ = hostsim.run(flips, [len(raw) + 64] * len(flips), exec_mode=3)
    for f, o, s in zip(flips, outs, status):
        ro, re_ = oracle_mod.decompress(f, len(raw) + 64)
        assert (re_ == 0) == (s == 0)
        if re_ == 0:
            assert o == ro
