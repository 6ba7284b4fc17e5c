"""GPU tests of the drop-in boundary (run with -m gpu): the reference's own ZstdTest cases
(kompressor-zstd--nativelib/src/commonTest/kotlin/com/ensody/kompressor/zstd/ZstdTest.kt) replayed
against ZstdDecompressor() of this build, plus the error behaviour of SURVEY.md Appendix C.
Frames are produced by the real libzstd (the compressor side stays on libzstd)."""
import base64

import numpy as np
import pytest

from helpers import gbytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(kzg_lib):
    from kompressor_b200 import core, io, zstd
    return core, io, zstd


def compress_like_kompressor(libzstd, data, level=3, dictionary=None):
    return libzstd.compress(data, level, dictionary=dictionary)


def test_tinySample(api, libzstd):                      # ZstdTest.kt:19-25
    core, _, zstd = api
    tiny = np.random.default_rng(1).integers(0, 256, 9000, dtype=np.uint8).tobytes()
    compressed = compress_like_kompressor(libzstd, tiny)
    assert core.transform(zstd.ZstdDecompressor(), compressed) == tiny


def test_sampleRoundtrip(api, libzstd):                 # ZstdTest.kt:27-32
    core, _, zstd = api
    data = np.random.default_rng(2).integers(0, 256, 1024 * 1024 + 3, dtype=np.uint8).tobytes()
    assert core.transform(zstd.ZstdDecompressor(), compress_like_kompressor(libzstd, data)) == data


def test_sampleRoundtripKotlinxIoSource(api, libzstd):  # ZstdTest.kt:34-38: pipe() shape, streamed frame
    _, io, zstd = api
    data = np.random.default_rng(3).integers(0, 256, 1024 * 1024 + 3, dtype=np.uint8).tobytes()
    compressed = libzstd.compress(data, 3, pledged=False, chunk=8192)
    assert io.pipe(compressed, zstd.ZstdDecompressor()).readByteArray() == data


def test_compressible_text_through_both_paths(api, libzstd):
    import corpus
    core, io, zstd = api
    data = bytes(corpus.generate("logs", 5, 700000))
    for pledged in (True, False):
        compressed = libzstd.compress(data, 3, checksum=True, pledged=pledged)
        assert core.transform(zstd.ZstdDecompressor(), compressed) == data
        assert io.pipe(compressed, zstd.ZstdDecompressor()).readByteArray() == data


def test_dictionaryRoundtrip(api, libzstd):             # ZstdTest.kt:49-65
    core, _, zstd = api
    dictionary = b"this is a test dictionary"
    data = b"this is some data to compress using a dictionary"
    compressed = compress_like_kompressor(libzstd, data, dictionary=dictionary)
    assert core.transform(zstd.ZstdDecompressor(dictionary=dictionary), compressed) == data
    with pytest.raises(zstd.IllegalStateException):     # must fail without the dictionary
        core.transform(zstd.ZstdDecompressor(), compressed)
    assert len(compressed) < len(compress_like_kompressor(libzstd, data))


def test_formattedDictionaryRoundtrip(api, libzstd):    # same constructor signature, ZDICT-trained dictionary
    import corpus
    core, io, zstd = api
    dictionary = libzstd.train_dictionary([bytes(corpus.generate("logs", 300 + i, 1500)) for i in range(200)], 8192)
    data = bytes(corpus.generate("logs", 777, 50000))
    compressed = compress_like_kompressor(libzstd, data, dictionary=dictionary)
    assert core.transform(zstd.ZstdDecompressor(dictionary=dictionary), compressed) == data
    assert io.pipe(compressed, zstd.ZstdDecompressor(dictionary=dictionary)).readByteArray() == data
    with pytest.raises(zstd.IllegalStateException):     # must fail without the dictionary
        core.transform(zstd.ZstdDecompressor(), compressed)


def test_sampleHello(api):                              # ZstdTest.kt:84-91
    core, _, zstd = api
    compressed = base64.b64decode("KLUv/QRYuQAAaGVsbG8gY29tcHJlc3Npb24gd29ybGR8Qm9f")
    assert core.transform(zstd.ZstdDecompressor(), compressed).decode() == "hello compression world"


def test_error_behaviour_appendix_c(api, manifest):
    core, _, zstd = api
    for e in manifest["errors"]:
        data = gbytes(e["name"] + ".zst")
        if e["error"]:
            with pytest.raises(zstd.IllegalStateException) as ei:
                core.transform(zstd.ZstdDecompressor(), data)
            assert ei.value.code == e["error"], e["name"]
            assert f"Bad zstd result code -{e['error']}:" in str(ei.value)
        else:
            out = core.transform(zstd.ZstdDecompressor(), data)   # truncated input: short output, no error
            assert len(out) <= e["out_size"]


def test_concatenated_and_skippable_frames(api, manifest):
    core, _, zstd = api
    for name in ("two_frames", "skippable_then_hello", "empty"):
        assert core.transform(zstd.ZstdDecompressor(), gbytes(name + ".zst")) == gbytes(name + ".raw")
    assert core.transform(zstd.ZstdDecompressor(), b"") == b""


def test_tiny_output_slices_and_byte_at_a_time_input(api, libzstd):
    """The streaming shim is resumable at arbitrary input and output cut points (64-byte output
    slices, SURVEY.md section 5 'Checkpoint / resume')."""
    import corpus
    core, _, zstd = api
    data = bytes(corpus.generate("json", 8, 50000))
    compressed = libzstd.compress(data, 3, checksum=True)
    d = zstd.ZstdDecompressor()
    out = bytearray()
    pos = 0
    step = 997
    while pos < len(compressed) or True:
        chunk = compressed[pos:pos + step]
        pos += len(chunk)
        inp = core.ByteArraySlice(bytearray(chunk))
        while True:
            o = core.ByteArraySlice(64)
            d.transform(inp, o, False)
            out += o.data[o.readStart:o.writeStart]
            if not (inp.hasData or o.insufficient):
                break
        if pos >= len(compressed):
            break
    assert bytes(out) == data
