"""GPU tests of the batch frame writer (run with -m gpu): every frame the encoder
emits must decode to the original bytes with the real libzstd (the decoder Kompressor's users have) and with this
repository's GPU decoder.  The reference's own round-trip tests (ZstdTest.kt:19-47) are the model."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def enc(kzg_lib):
    from kompressor_b200.batch import BatchEncoder
    e = BatchEncoder()
    yield e
    e.close()


def roundtrip(enc, libzstd, data: bytes, chunk=65536, checksum=False):
    comp = enc.compress(data, chunk_size=chunk, checksum=checksum)
    out, err = libzstd.decompress_oneshot(comp, len(data) + 64)
    assert err == 0, err
    assert out == data
    return comp


def test_corpora_roundtrip_through_libzstd(enc, libzstd):
    import corpus
    for kind in ("logs", "json", "literal"):
        data = bytes(corpus.generate(kind, 31, 1 << 20))
        for chunk, cs in ((65536, False), (131072, True), (16384, True), (5000, False)):
            comp = roundtrip(enc, libzstd, data, chunk, cs)
            if kind != "literal" and chunk >= 65536:
                assert len(comp) < 0.6 * len(data), (kind, chunk, len(comp))


def test_edge_cases(enc, libzstd):
    rng = np.random.default_rng(5)
    cases = [b"", b"a", b"ab" * 31, b"x" * 64, b"x" * 65, b"q" * 131072, b"abcdefgh" * 16384,
             rng.integers(0, 256, 65536, dtype=np.uint8).tobytes(), rng.integers(0, 2, 100000, dtype=np.uint8).tobytes(),
             bytes(range(256)) * 300, b"hello compression world"]
    for data in cases:
        for cs in (False, True):
            roundtrip(enc, libzstd, data, 65536, cs)
            roundtrip(enc, libzstd, data, 131072, cs)
    # empty input gives the same frame libzstd writes (golden vector 3)
    assert enc.compress(b"").hex() == "28b52ffd2000010000"


def test_status_codes(enc):
    data = np.zeros(300000, np.uint8)
    # a chunk above 128 KiB is refused (72); the others still succeed
    offs = np.array([0, 200000], np.uint64)
    sizes = np.array([200000, 1000], np.uint64)
    out, doff, osz, status = enc.compress_chunks(data, offs, sizes)
    assert status[0] == 72 and status[1] == 0 and osz[1] > 0


def test_gpu_decoder_reads_gpu_frames(enc, libzstd):
    import corpus
    from kompressor_b200.mirror import core, zstd
    data = bytes(corpus.generate("logs", 77, 3 << 20)) + bytes(corpus.generate("json", 78, 1 << 20))
    comp = enc.compress(data, checksum=True)
    assert core.transform(zstd.ZstdDecompressor(), comp) == data


def test_output_is_deterministic(enc):
    import corpus
    data = bytes(corpus.generate("logs", 9, 1 << 20))
    assert enc.compress(data) == enc.compress(data)


def test_level_selects_the_parse(enc, libzstd):
    """compressionLevel below 3: greedy parse; from 3 (Kompressor's default): one step of lazy matching, never larger
    on text.  Both decode with libzstd."""
    import corpus
    data = bytes(corpus.generate("logs", 55, 2 << 20))
    greedy = enc.compress(data, level=1)
    lazy = enc.compress(data, level=3)
    for comp in (greedy, lazy):
        out, err = libzstd.decompress_oneshot(comp, len(data) + 64)
        assert err == 0 and out == data
    assert len(lazy) < len(greedy)


def test_seek_table_stream(enc, libzstd):
    """The frame writer's stream with a seekable index at its end decodes everywhere, and the index gives the frame
    table without a walk."""
    import corpus
    from kompressor_b200.mirror import core, zstd
    from kompressor_b200.batch import scan_frames
    data = bytes(corpus.generate("json", 91, 1000003))
    comp = enc.compress(data, seek_table=True)
    assert comp[-4:] == bytes.fromhex("b1ea928f")
    off, size, content, consumed = scan_frames(comp)
    assert consumed == len(comp) and len(off) == 16 + 1 and int(content[:16].sum()) == len(data)
    out, err = libzstd.decompress_oneshot(comp, len(data) + 64)
    assert err == 0 and out == data
    assert core.transform(zstd.ZstdDecompressor(), comp) == data
