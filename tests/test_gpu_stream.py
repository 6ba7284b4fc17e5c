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


def test_many_handles_in_parallel(api, libzstd):
    """One handle per thread, as one per HTTP body (SURVEY.md section 8(b) 'Threading'): the shim merges the
    frames of concurrent calls into shared GPU batches; every stream must still see exactly its own bytes,
    its own dictionary and its own errors."""
    import threading
    import corpus
    core, io, zstd = api
    dictionary = libzstd.train_dictionary([bytes(corpus.generate("logs", 300 + i, 1500)) for i in range(200)], 8192)
    raw_dict = b"level=INFO component=scheduler thread=worker-"
    jobs = []
    for t in range(24):
        kind = ("logs", "json", "literal")[t % 3]
        frames = [bytes(corpus.generate(kind, 1000 + 17 * t + i, 20000 + 3001 * ((t + i) % 5))) for i in range(6)]
        d = (None, dictionary, raw_dict)[t % 4] if t % 4 < 3 else None
        comp = b"".join(libzstd.compress(f, 3, checksum=bool(t & 1), dictionary=d) for f in frames)
        jobs.append(dict(data=b"".join(frames), comp=comp, dict=d, mode="pipe" if t % 2 else "transform", bad=False))
    big = bytes(corpus.generate("logs", 4242, 24 << 20))          # a large request that runs alone between the merged ones
    jobs.append(dict(data=big, comp=b"".join(libzstd.compress(big[i:i + (1 << 20)], 3) for i in range(0, len(big), 1 << 20)),
                     dict=None, mode="transform", bad=False))
    broken = bytearray(b"".join(libzstd.compress(bytes(corpus.generate("logs", 9000 + i, 30000)), 3, checksum=True) for i in range(4)))
    broken[200] ^= 0x55                                          # inside the first frame's literals: corruption or checksum error
    jobs.append(dict(data=None, comp=bytes(broken), dict=None, mode="pipe", bad=True))
    results = [None] * len(jobs)

    def work(i):
        j = jobs[i]
        try:
            for _ in range(3):
                dec = zstd.ZstdDecompressor(dictionary=j["dict"])
                out = core.transform(dec, j["comp"]) if j["mode"] == "transform" else io.pipe(j["comp"], dec).readByteArray()
                dec.close()
                if out != j["data"]:
                    results[i] = "mismatch"
                    return
            results[i] = "ok"
        except zstd.IllegalStateException as e:
            results[i] = f"error {e.code}"

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for i, j in enumerate(jobs):
        if j["bad"]:
            assert results[i] is not None and results[i].startswith("error"), (i, results[i])
        else:
            assert results[i] == "ok", (i, results[i])


# ---------------------------------------------------------------- the reference's round trips with this build's compressor
def test_roundtrip_with_own_compressor(api, libzstd):   # ZstdTest.kt:19-38 (tinySample, sampleRoundtrip, ...KotlinxIoSource)
    import corpus
    core, io, zstd = api
    tiny = np.random.default_rng(11).integers(0, 256, 9000, dtype=np.uint8).tobytes()
    text = bytes(corpus.generate("logs", 12, 1024 * 1024 + 3))
    rnd = np.random.default_rng(13).integers(0, 256, 1024 * 1024 + 3, dtype=np.uint8).tobytes()
    for data in (tiny, text, rnd, b"", b"x"):
        compressed = core.transform(zstd.ZstdCompressor(compressionLevel=3), data)
        assert core.transform(zstd.ZstdDecompressor(), compressed) == data
        out, err = libzstd.decompress_oneshot(compressed, len(data) + 64)       # and the real libzstd reads it
        assert err == 0 and out == data
    assert len(core.transform(zstd.ZstdCompressor(), text)) < len(text) // 2
    source = io.pipe(io.pipe(text, zstd.ZstdCompressor(compressionLevel=3)), zstd.ZstdDecompressor())
    assert source.readByteArray() == text


def test_compressor_parameters_and_dictionaries(api, libzstd):
    core, _, zstd = api
    data = b"this is some data to compress using a dictionary" * 50
    c = zstd.ZstdCompressor()
    c.setParameter(201, 1)                               # ZSTD_c_checksumFlag
    framed = core.transform(c, data)
    assert framed[4] & 4                                 # Content_Checksum flag of the frame header descriptor
    assert core.transform(zstd.ZstdDecompressor(), framed) == data
    with pytest.raises(zstd.IllegalStateException) as ei:
        c.setParameter(9999, 1)
    assert ei.value.code == 40
    # a raw-content dictionary is accepted (the frames do not reference it, so they decode with or without it)
    raw_dict = b"this is a test dictionary"
    compressed = core.transform(zstd.ZstdCompressor(dictionary=raw_dict), data)
    assert core.transform(zstd.ZstdDecompressor(dictionary=raw_dict), compressed) == data
    assert core.transform(zstd.ZstdDecompressor(), compressed) == data
    # a formatted dictionary is refused instead of silently producing frames that need other repeat offsets
    import corpus
    trained = libzstd.train_dictionary([bytes(corpus.generate("logs", 300 + i, 1500)) for i in range(200)], 8192)
    with pytest.raises(zstd.IllegalStateException) as ei:
        zstd.ZstdCompressor(dictionary=trained)
    assert ei.value.code == 40


def test_sampleRoundtripKotlinxIoSink(api, libzstd):    # ZstdTest.kt:40-47: compressor.pipe(decompressor.pipe(output))
    import corpus
    _, io, zstd = api
    for data in (np.random.default_rng(21).integers(0, 256, 1024 * 1024 + 3, dtype=np.uint8).tobytes(),
                 bytes(corpus.generate("json", 22, 700001))):
        output = io.BufferSink()
        sink = io.pipe(zstd.ZstdCompressor(compressionLevel=3), io.pipe(zstd.ZstdDecompressor(), output))
        io.transferTo(data, sink)
        sink.close()
        assert output.closed and output.readByteArray() == data
        # the decompressor alone behind a sink, fed with libzstd's frame
        output = io.BufferSink()
        sink = io.pipe(zstd.ZstdDecompressor(), output)
        io.transferTo(libzstd.compress(data, 3, checksum=True), sink)
        sink.close()
        assert output.readByteArray() == data


def test_frames_before_a_bad_frame_are_delivered_first(api, libzstd):
    """libzstd hands over the output of the frames in front of a corrupt one before it reports the error:
    the shim decodes all complete frames of a call as one batch and must behave the same."""
    import corpus
    core, _, zstd = api
    raws = [bytes(corpus.generate("logs", 600 + i, 30000)) for i in range(4)]
    frames = [bytearray(libzstd.compress(r, 3, checksum=True)) for r in raws]
    frames[2][len(frames[2]) // 2] ^= 0x10
    blob = b"".join(bytes(f) for f in frames)
    d = zstd.ZstdDecompressor()
    inp = core.ByteArraySlice(bytearray(blob))
    inp.writeStart = len(blob)
    got = bytearray()
    with pytest.raises(zstd.IllegalStateException) as ei:
        while True:
            o = core.ByteArraySlice(8192)
            d.transform(inp, o, True)
            got += o.data[o.readStart:o.writeStart]
            if not (inp.hasData or o.insufficient):
                break
    assert ei.value.code in (20, 22)
    assert bytes(got) == raws[0] + raws[1]
