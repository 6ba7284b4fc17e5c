"""GPU parity (run with -m gpu on a B200): the CUDA path, called through the C ABI (libkzg.so), must
be bit-exact against the committed golden fixtures, the oracle and the real libzstd on the same
seeded inputs.  Integer/byte work: the bar is bit-exact, there is no tolerance."""
import base64
import hashlib

import numpy as np
import pytest

from helpers import check_output, frames_of, gbytes, table_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dec(kzg_lib):
    from kompressor_b200.batch import BatchDecoder
    d = BatchDecoder()
    yield d
    d.close()


def run(dec, frames, caps):
    comp, offs, sizes = table_of(frames)
    out, doff, status, osz = dec.decompress_host(comp, offs, sizes, np.array(caps, dtype=np.uint64))
    return [bytes(out[int(doff[i]):int(doff[i] + osz[i])]) for i in range(len(frames))], status, osz


def test_golden_vectors_one_batch(dec, manifest):
    vs = [v for v in manifest["vectors"] if "dictionary_b64" not in v and v["name"] not in ("skippable_then_hello", "two_frames")]
    outs, status, _ = run(dec, [gbytes(v["name"] + ".zst") for v in vs], [v["size"] + 32 for v in vs])
    for v, o, s in zip(vs, outs, status):
        assert s == 0, v["name"]
        check_output(v, o)


def test_golden_error_vectors(dec, manifest):
    # codes libzstd returns for the same frames (one-shot decode reports truncation as srcSize_wrong)
    expect = {e["name"]: e["error"] for e in manifest["errors"]}
    expect["err_truncated"] = 72
    expect["err_missing_checksum_byte"] = 72
    del expect["err_trailing_garbage"]           # entry = exactly one frame; covered by the streaming tests
    names = sorted(expect)
    _, status, _ = run(dec, [gbytes(n + ".zst") for n in names], [1 << 16] * len(names))
    for n, s in zip(names, status):
        assert s == expect[n], (n, s)


def test_documented_divergences(dec, manifest):
    """Where this decoder deliberately differs from libzstd: a Huffman stream that runs out of bits in
    a frame without checksum -- libzstd makes up literals and accepts, the GPU path reports corruption_detected."""
    ds = manifest["divergences"]
    _, status, _ = run(dec, [gbytes(d["name"] + ".zst") for d in ds], [1 << 16] * len(ds))
    for d, s in zip(ds, status):
        assert s == d["ours"], (d["name"], s)


@pytest.mark.parametrize("kind,level,checksum", [("logs", 3, False), ("json", 3, True), ("literal", 1, False), ("random", 3, False)])
def test_corpora_against_oracle_and_libzstd(dec, oracle_mod, libzstd, kind, level, checksum):
    import corpus
    c = corpus.build(kind, 96, 65536, level, checksum)
    frames = frames_of(c)
    outs, status, sizes = run(dec, frames, [65536] * c.n_frames)
    assert (status == 0).all() and (sizes == 65536).all()
    assert b"".join(outs) == c.raw.tobytes()
    for i in (0, 17, 95):
        o, e = oracle_mod.decompress(frames[i], 65536)
        r, re_ = libzstd.decompress_oneshot(frames[i], 65536)
        assert e == 0 and re_ == 0 and outs[i] == o == r


def test_readme_roundtrip_config(dec, libzstd):
    # BASELINE.json configs[0]: 128 KiB + 3 random bytes, level 3 -> raw blocks 131072 + 3
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 256, 128 * 1024 + 3, dtype=np.uint8).tobytes()
    fr = libzstd.compress(raw, 3)
    assert len(fr) == 131090
    outs, status, _ = run(dec, [fr], [len(raw)])
    assert status[0] == 0 and outs[0] == raw


def test_ragged_sizes_and_empty(dec, libzstd):
    import corpus
    sizes = [0, 1, 2, 3, 5, 6, 7, 31, 32, 33, 255, 256, 1000, 4097, 65535, 65536, 65537, 131072, 200001]
    raws = [bytes(corpus.generate("json", 40 + i, n)) for i, n in enumerate(sizes)]
    frames = [libzstd.compress(r, 3, checksum=bool(i & 1)) for i, r in enumerate(raws)]
    outs, status, _ = run(dec, frames, [len(r) for r in raws])
    assert (status == 0).all()
    assert outs == raws
    # one byte of room too few -> dstSize_tooSmall (70), never a write past the capacity
    outs, status, _ = run(dec, frames[4:], [len(r) - 1 for r in raws[4:]])
    assert (status == 70).all()


def test_multiblock_levels_and_streamed_headers(dec, libzstd):
    import corpus
    raw = bytes(corpus.generate("mixed", 77, 900000))
    frames = [libzstd.compress(raw, lvl, checksum=True, pledged=p) for lvl in (-5, 1, 3, 9, 19) for p in (True, False)]
    outs, status, _ = run(dec, frames, [len(raw)] * len(frames))
    assert (status == 0).all()
    for o in outs:
        assert hashlib.sha256(o).digest() == hashlib.sha256(raw).digest()


def test_run_like_data_overlapping_matches(dec, libzstd):
    rng = np.random.default_rng(5)
    raws = [bytes(70000), b"ab" * 40000, b"abc" * 30000, bytes(rng.integers(0, 4, 100000, dtype=np.uint8)),
            (b"0123456789abcdef" * 7 + b"x") * 900]
    frames = [libzstd.compress(r, lvl) for r in raws for lvl in (1, 3, 19)]
    outs, status, _ = run(dec, frames, [len(r) for r in raws for _ in range(3)])
    assert (status == 0).all()
    assert outs == [r for r in raws for _ in range(3)]


def test_far_offsets_and_wide_sequences(dec, libzstd):
    """8 MiB window (level 19, streamed header): matches millions of bytes back, literal runs and matches
    of 64 KiB+ in one sequence (more than 32 extra bits per sequence: the sequence kernel's general path)."""
    rng = np.random.default_rng(21)
    a = rng.integers(0, 256, 5 << 20, dtype=np.uint8).tobytes()
    b = rng.integers(0, 256, 70000, dtype=np.uint8).tobytes()
    raw = a + b + a[1000:121000] + b[:3] + a[4 << 20:(4 << 20) + 70000] + b + a[77:131149]
    frames = [libzstd.compress(raw, 19, checksum=True, pledged=p) for p in (True, False)]
    outs, status, _ = run(dec, frames, [len(raw)] * len(frames))
    assert (status == 0).all()
    for o in outs:
        assert hashlib.sha256(o).digest() == hashlib.sha256(raw).digest()


def test_formatted_dictionary_batch(dec, oracle_mod, libzstd, manifest):
    """Formatted dictionary (ZDICT-trained: entropy tables, repeat offsets, content, dictID) on the batch layer:
    the golden vector, frames of several sizes incl. multi-block ones, the wrong dictionary, no dictionary."""
    import base64
    import corpus
    from helpers import gbytes as gb
    v = next(x for x in manifest["vectors"] if x["name"] == "dict_formatted_json_3k")
    fdict = base64.b64decode(v["dictionary_b64"])
    raws = [bytes(corpus.generate("json", 500 + i, n)) for i, n in enumerate((0, 1, 40, 700, 3000, 70000, 300000))]
    frames = [gb(v["name"] + ".zst")] + [libzstd.compress(r, lvl, dictionary=fdict, checksum=True) for r in raws for lvl in (1, 3, 19)]
    expect = [gb(v["name"] + ".raw")] + [r for r in raws for _ in range(3)]
    try:
        dec.set_dictionary(fdict)
        outs, status, _ = run(dec, frames, [len(r) + 8 for r in expect])
        assert (status == 0).all(), status
        assert outs == expect
        for f, e in list(zip(frames, expect))[::5]:
            o, err = oracle_mod.decompress(f, len(e) + 8, fdict)
            assert err == 0 and o == e
        # a dictionary with another ID: frames name their dictID -> dictionary_wrong (32), as libzstd
        other = bytearray(fdict); other[4] ^= 0x5A
        dec.set_dictionary(bytes(other))
        _, status, _ = run(dec, frames[:4], [len(r) + 8 for r in expect[:4]])
        assert (status == 32).all(), status
        _, lerr = libzstd.decompress_oneshot(frames[0], 65536, dictionary=bytes(other))
        assert lerr == 32
    finally:
        dec.set_dictionary(None)
    _, status, _ = run(dec, frames[:4], [len(r) + 8 for r in expect[:4]])
    assert (status != 0).all()


def test_unaligned_output_offsets(dec, libzstd):
    import corpus
    raws = [bytes(corpus.generate("logs", 90 + i, 30000 + 17 * i)) for i in range(9)]
    frames = [libzstd.compress(r, 3) for r in raws]
    comp, offs, sizes = table_of(frames)
    caps = np.array([len(r) for r in raws], dtype=np.uint64)
    doff = np.zeros(len(raws), np.uint64)
    doff[1:] = np.cumsum(caps + np.arange(1, len(raws) + 1, dtype=np.uint64))[:-1]     # odd gaps -> odd alignments
    out = np.full(int(doff[-1] + caps[-1]) + 64, 0xEE, np.uint8)
    out, _, status, osz = dec.decompress_host(comp, offs, sizes, caps, out=out, out_offsets=doff)
    assert (status == 0).all()
    for i, r in enumerate(raws):
        assert bytes(out[int(doff[i]):int(doff[i]) + len(r)]) == r


def test_bitflips_agree_with_oracle(dec, oracle_mod, libzstd):
    """Corrupt frames: the CUDA path must make the same accept/reject decision as the oracle, and the
    same bytes when both accept."""
    import corpus
    rng = np.random.default_rng(11)
    raw = bytes(corpus.generate("json", 11, 30000))
    fr = libzstd.compress(raw, 3, checksum=False)
    frames = []
    for _ in range(256):
        b = bytearray(fr)
        b[int(rng.integers(4, len(b)))] ^= 1 << int(rng.integers(0, 8))
        frames.append(bytes(b))
    outs, status, _ = run(dec, frames, [65536] * len(frames))
    for f, o, s in zip(frames, outs, status):
        ro, re_ = oracle_mod.decompress(f, 65536)
        assert (re_ == 0) == (s == 0), (re_, s)
        if re_ == 0:
            assert o == ro


def test_device_resident_batch_and_size_independent_properties(dec):
    """Device-pointer form at a larger size: every frame decodes to its exact size, decode is
    idempotent, and a checksum of per-frame checksums matches the generator's."""
    import torch
    import corpus
    n = 2048
    c = corpus.build("logs", n, 65536, 3, False)
    caps = np.full(n, 65536, np.uint64)
    b = dec.upload(c.comp, c.offsets, c.sizes, caps)
    dec.run(b)
    assert bool((b.status == 0).all()) and bool((b.out_size == 65536).all())
    first = b.dst[:n * 65536].clone()
    dec.run(b)
    assert torch.equal(first, b.dst[:n * 65536])
    got = first.cpu().numpy().reshape(n, 65536)
    exp = c.raw.reshape(n, 65536)
    h_got = hashlib.sha256(b"".join(hashlib.sha256(r.tobytes()).digest() for r in got)).hexdigest()
    h_exp = hashlib.sha256(b"".join(hashlib.sha256(r.tobytes()).digest() for r in exp)).hexdigest()
    assert h_got == h_exp
    st = dec.stats()
    assert st.n_cblocks == n and st.launches in (11, 13)      # the Huffman stage is two launches, the execute stage one or three


def test_hostile_content_size_does_not_reach_the_allocator(dec, libzstd):
    """a 17-byte frame claiming 2^64 - 2 bytes in front of a valid 1 MiB frame.  With the capacities
    kzg_scan_frames hands out, the hostile frame fails (corruption_detected, as in libzstd) and the valid one decodes;
    capacities that wrap are refused by the host entry point instead of reaching device memory."""
    import struct
    from kompressor_b200 import KzgError
    from kompressor_b200.batch import scan_frames
    import corpus
    hostile = struct.pack("<IB", 0xFD2FB528, 0xE0) + struct.pack("<Q", 0xFFFFFFFFFFFFFFFE) + bytes([0x01, 0x00, 0x00])
    raw = bytes(corpus.generate("logs", 31, 1 << 20))
    good = libzstd.compress(raw, 3)
    blob = hostile + good
    off, size, content, _ = scan_frames(blob)
    comp = np.frombuffer(blob, np.uint8).copy()
    out, doff, status, osz = dec.decompress_host(comp, off, size, content)
    assert int(status[0]) == 20 and int(status[1]) == 0
    assert bytes(out[int(doff[1]):int(doff[1] + osz[1])]) == raw
    with pytest.raises(KzgError):       # the wrapping capacity itself: refused before any device arithmetic
        dec.decompress_host(comp, off, size, np.array([0xFFFFFFFFFFFFFFF0, len(raw)], np.uint64),
                            out=np.zeros(len(raw) + 64, np.uint8), out_offsets=np.array([0, 0], np.uint64))
    # the context is still usable afterwards (a failed sizing must not stick)
    outs, status, _ = run(dec, [good], [len(raw)])
    assert status[0] == 0 and outs[0] == raw


@pytest.mark.parametrize("tile", ["4096", "8192"])
def test_second_execute_design_is_bit_exact(dec, libzstd, tile, monkeypatch):
    """kzg_exec2.cuh (output-centric execute, ring window, bulk flush) is selectable with KZG_EXEC=2; it must give the
    same bytes as the default kernel on every corpus kind, on runs, long far matches and multi-block level-19 frames."""
    import corpus
    monkeypatch.setenv("KZG_EXEC", "2")
    monkeypatch.setenv("KZG_EXEC_TILE", tile)
    rng = np.random.default_rng(3)
    raws = []
    for kind in ("logs", "json", "literal"):
        raws += [bytes(corpus.generate(kind, 50 + i, 65536)) for i in range(24)]
    raws += [bytes(40000), b"abc" * 20000, bytes(rng.integers(0, 3, 70000, dtype=np.uint8))]
    a = bytes(corpus.generate("logs", 9, 50000))
    raws.append(a + a[::-1] + a)
    raws += [bytes(corpus.generate("mixed", 70 + i, 700000)) for i in range(2)]
    frames = [libzstd.compress(r, 19 if len(r) > 200000 else 3, checksum=bool(i & 1), pledged=bool(i & 2)) for i, r in enumerate(raws)]
    outs, status, _ = run(dec, frames, [len(r) + (i % 5) for i, r in enumerate(raws)])
    assert (status == 0).all()
    assert outs == raws


def test_configs4_exact_shape(dec, libzstd):
    """BASELINE.json configs[4]: one 4 MiB frame, 2 MiB log text + 2 MiB JSON lines, level 19, streamed (window
    descriptor, 8 MiB window, no content size), cross-block history and long offsets."""
    import corpus
    c = corpus.build("mixed", 2, 4 << 20, 19, False, streamed=True)
    frames = frames_of(c)
    assert frames[0][4] & 0x20 == 0                       # not single-segment: a window descriptor follows
    outs, status, sizes = run(dec, frames, [4 << 20] * 2)
    assert (status == 0).all() and (sizes == 4 << 20).all()
    assert hashlib.sha256(b"".join(outs)).digest() == hashlib.sha256(c.raw.tobytes()).digest()


def test_first_design_window_paths(dec, libzstd, monkeypatch):
    """The warp-per-frame execute kernel (kzg_exec.cuh, the batch path) on inputs made for the paths of its shared-memory
    window -- small batches normally go to the CTA-per-frame kernel, KZG_EXEC_CTA_BELOW=0 sends them here:
    literal runs of 17 .. 64 bytes (the copy loop instead of the straight-line placement), of more than 64 and more
    than 512 bytes (warp-wide copies, the literal ring skipping ahead), long matches in the middle of a batch (the
    window starts over and the sources of the remaining lanes are staged again), short periods, sources right at the
    window's lower edge, and frames of up to eight blocks."""
    import corpus
    monkeypatch.setenv("KZG_EXEC_CTA_BELOW", "0")
    rng = np.random.default_rng(17)
    text = bytes(corpus.generate("logs", 123, 400000))
    js = bytes(corpus.generate("json", 321, 200000))

    def noise(n):
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()

    raws = []
    # literal runs of every length class between short matches
    for lo, hi in ((1, 16), (17, 64), (65, 300), (400, 2000), (1, 3000)):
        parts, pos = [], 0
        while sum(map(len, parts)) < 120000:
            parts.append(noise(int(rng.integers(lo, hi + 1))))
            n = int(rng.integers(4, 40))
            parts.append(text[pos:pos + n]); pos = (pos + 977) % 300000
        raws.append(b"".join(parts))
    # long matches (65 .. 5000 bytes) from near and far in the middle of ordinary text
    for span in (3000, 60000):
        parts, pos = [text[:span]], span
        while sum(map(len, parts)) < 200000:
            parts.append(text[pos:pos + int(rng.integers(20, 400))]); pos += 400
            back = int(rng.integers(100, span))
            hist = b"".join(parts)
            parts.append(hist[-back:][:int(rng.integers(65, 5000))])
        raws.append(b"".join(parts))
    # short periods and runs between text, sources a few hundred bytes back (the window's lower edge after a slide)
    parts = []
    for i in range(400):
        parts.append(text[i * 300:i * 300 + int(rng.integers(30, 200))])
        parts.append(bytes([65 + i % 7]) * int(rng.integers(1, 90)))
        parts.append((b"xy", b"abc", b"12345", b"abcdefg")[i % 4] * int(rng.integers(1, 30)))
    raws.append(b"".join(parts))
    raws.append(b"".join(text[i:i + 37] + text[max(0, i - 390):max(0, i - 390) + 21] + text[max(0, i - 3200):max(0, i - 3200) + 33]
                         for i in range(4000, 160000, 180)))
    # plain corpora in single-block and multi-block frames (at most eight blocks stay with this kernel)
    raws += [text[:65536], js[:65536], text[:131072 * 3 + 17], (js + text)[:131072 * 8], bytes(corpus.generate("mixed", 5, 500000))]
    frames, caps, expect = [], [], []
    for i, r in enumerate(raws):
        for lvl in (1, 3, 9, 19):
            frames.append(libzstd.compress(r, lvl, checksum=bool(i & 1), pledged=bool(lvl & 1)))
            caps.append(len(r) + (i % 3))
            expect.append(r)
    outs, status, _ = run(dec, frames, caps)
    assert (status == 0).all(), status
    for k, (o, e) in enumerate(zip(outs, expect)):
        assert o == e, k
