"""The CUDA decoder's per-thread stage code (kzg_stages.cuh: scan, parse, link, Huffman, FSE, offsets,
packed sequence records, symbolic repeat offsets) compiled for the host and checked against the golden
fixtures and the oracle.  This is host logic testing only -- the GPU parity tests are in test_gpu_*.py."""
import base64
import hashlib

import numpy as np
import pytest

from helpers import check_output, frames_of, gbytes

hostsim = pytest.importorskip("hostsim")


def test_golden_vectors(manifest):
    vs = [v for v in manifest["vectors"] if "dictionary_b64" not in v and v["name"] not in ("skippable_then_hello", "two_frames")]
    outs, status, _ = hostsim.run([gbytes(v["name"] + ".zst") for v in vs], [v["size"] + 32 for v in vs])
    for v, o, s in zip(vs, outs, status):
        assert s == 0, v["name"]
        check_output(v, o)


def test_error_vectors(manifest):
    expect = {"err_checksum_flipped": 22, "err_magic": 10, "err_corrupt_block": 22, "err_reserved_block": 20,
              "err_fcs_mismatch": 20, "err_windowlog_28": 16, "err_needs_dictionary": 20, "err_reserved_bit": 14,
              "err_truncated": 72, "err_missing_checksum_byte": 72}
    names = sorted(expect)
    outs, status, _ = hostsim.run([gbytes(n + ".zst") for n in names], [1 << 16] * len(names))
    for n, s in zip(names, status):
        assert s == expect[n], (n, s)


@pytest.mark.parametrize("kind,level,checksum", [("logs", 3, False), ("json", 3, True), ("literal", 1, False), ("random", 3, False)])
def test_corpora(kind, level, checksum):
    import corpus
    c = corpus.build(kind, 16, 65536, level, checksum)
    outs, status, sizes = hostsim.run(frames_of(c), [65536] * c.n_frames)
    assert (status == 0).all() and (sizes == 65536).all()
    assert b"".join(outs) == c.raw.tobytes()


def test_multiblock_symbolic_repeat_offsets(libzstd):
    import corpus
    raw = bytes(corpus.generate("mixed", 77, 700000))
    frames = [libzstd.compress(raw, lvl, checksum=True, pledged=p) for lvl in (3, 19) for p in (True, False)]
    outs, status, _ = hostsim.run(frames, [len(raw)] * len(frames))
    assert (status == 0).all()
    for o in outs:
        assert hashlib.sha256(o).digest() == hashlib.sha256(raw).digest()


def test_dst_too_small_and_ragged_sizes(libzstd):
    import corpus
    raws = [bytes(corpus.generate("logs", 5 + i, n)) for i, n in enumerate([0, 1, 2, 3, 7, 100, 1000, 4097, 65535, 65537, 200001])]
    frames = [libzstd.compress(r, 3) for r in raws]
    outs, status, _ = hostsim.run(frames, [len(r) for r in raws])
    assert (status == 0).all() and outs == raws
    outs, status, _ = hostsim.run(frames[5:], [len(r) - 1 for r in raws[5:]])
    assert (status == 70).all()


def test_encode_stage_roundtrip_through_libzstd(libzstd):
    """Compress side (kzg_enc_stage_encode, the code the GPU's encode kernel runs): frames written behind a sequential
    stand-in matcher must decode with libzstd to the original bytes; empty input gives libzstd's own empty frame."""
    import corpus
    chunks = [bytes(corpus.generate(k, 100 + i, sz)) for i, (k, sz) in enumerate(
        [("logs", 65536), ("json", 65536), ("literal", 65536), ("logs", 131072), ("logs", 1000), ("json", 300), ("logs", 64),
         ("logs", 10), ("logs", 0)])]
    chunks += [bytes(np.random.default_rng(1).integers(0, 256, 65536, dtype=np.uint8)), b"a" * 70000, b"abcdefgh" * 9000]
    for checksum in (False, True):
        frames, status = hostsim.compress(chunks, checksum=checksum)
        for c, f, st in zip(chunks, frames, status):
            assert st == 0
            out, err = libzstd.decompress_oneshot(f, len(c) + 16)
            assert err == 0 and out == c
    assert hostsim.compress([b""])[0][0].hex() == "28b52ffd2000010000"
    assert len(hostsim.compress([chunks[0]])[0][0]) < 0.4 * 65536


def test_fse_coded_huffman_weights_roundtrip():
    """Compress side: the encoder's FSE-coded tree description (kzg_enc_compress_weights) read back by the decoder's
    parser (kzg_read_huf_weights) for random complete codes of up to 256 symbols and 11 bits."""
    rng = np.random.default_rng(3)
    checked = 0
    for _ in range(300):
        m = int(rng.integers(2, 257))
        leaves = [0]
        while len(leaves) < m:
            cands = [j for j, x in enumerate(leaves) if x < 11]
            if not cands:
                break
            i = cands[int(rng.integers(0, len(cands)))]
            leaves[i] += 1
            leaves.append(leaves[i])
        m = len(leaves)
        maxbits = max(leaves)
        nsym = int(rng.integers(m, 257))
        syms = np.unique(np.append(rng.choice(nsym - 1, m - 1, replace=False), nsym - 1)) if m > 1 else np.array([nsym - 1])
        if len(syms) != m:
            continue
        w = np.zeros(nsym, np.uint8)
        for s_, l in zip(syms, rng.permutation(np.array(leaves))):
            w[s_] = maxbits + 1 - l
        r = hostsim.weights_roundtrip(w[:-1])
        assert r >= 0
        checked += r > 0
    assert checked > 200
