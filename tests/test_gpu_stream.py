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
    from kompressor_b200.mirror import core, io, zstd
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
            assert len(out) == e["out_size"], e["name"]


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


# ------------------------------------------------------------------------------------------------
# Call-for-call behaviour against the real ZSTD_decompressStream (Wrapper.cpp:178): the same driver runs libzstd 1.5.7
# and the shim through ctypes with the reference's call patterns and compares what a caller can observe -- the bytes
# handed over by the calls that did not fail, and the error code.
import ctypes


class _InBuf(ctypes.Structure):
    _fields_ = [("src", ctypes.c_void_p), ("size", ctypes.c_size_t), ("pos", ctypes.c_size_t)]


class _OutBuf(ctypes.Structure):
    _fields_ = [("dst", ctypes.c_void_p), ("size", ctypes.c_size_t), ("pos", ctypes.c_size_t)]


def _drive(create, free, step, is_error, data, out_slice, in_chunk=None, dictionary=None, load_dict=None):
    """SliceTransform.transform(ByteArray) (in_chunk None: whole input visible, SliceTransform.kt:33-45) or a pipe()-like
    pump (in_chunk bytes become visible per round, SliceTransformRawSource.kt:32-55).  Output slices are reused until full.
    Returns (bytes handed over, error code, offsets at which a call returned 0)."""
    dctx = create()
    assert dctx
    try:
        if dictionary is not None:
            load_dict(dctx, dictionary, len(dictionary))
        src = ctypes.create_string_buffer(bytes(data), max(1, len(data)))
        base = ctypes.addressof(src)
        out, zeros = bytearray(), []
        visible, pos = (len(data) if in_chunk is None else 0), 0
        dst = ctypes.create_string_buffer(out_slice)
        wpos = 0
        for _ in range(10_000_000):
            if in_chunk is not None and pos == visible and visible < len(data):
                visible = min(len(data), visible + in_chunk)
            inb = _InBuf(base, visible, pos)
            outb = _OutBuf(ctypes.addressof(dst), out_slice, wpos)
            r = step(dctx, ctypes.byref(outb), ctypes.byref(inb))
            if is_error(r):
                return bytes(out), (1 << 64) - (r & ((1 << 64) - 1)), zeros
            out += dst.raw[wpos:outb.pos]
            pos, wpos = inb.pos, outb.pos
            if r == 0:
                zeros.append(len(out))
            insufficient = wpos == out_slice and r != 0
            if wpos == out_slice:
                wpos = 0
            if pos == visible and not insufficient and visible == len(data):
                return bytes(out), 0, zeros
        raise AssertionError("driver did not terminate")
    finally:
        free(dctx)


@pytest.fixture(scope="module")
def both(kzg_lib, libzstd):
    Z, K = libzstd.lib, kzg_lib
    for L, pre in ((Z, "ZSTD_"), (K, "kzg_")):
        getattr(L, pre + "createDCtx").restype = ctypes.c_void_p
        getattr(L, pre + "freeDCtx").argtypes = [ctypes.c_void_p]
        getattr(L, pre + "decompressStream").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        getattr(L, pre + "decompressStream").restype = ctypes.c_size_t
        getattr(L, pre + "DCtx_loadDictionary").argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
        getattr(L, pre + "DCtx_loadDictionary").restype = ctypes.c_size_t
        getattr(L, pre + "isError").argtypes = [ctypes.c_size_t]

    def run(L, pre, data, out_slice, in_chunk=None, dictionary=None):
        return _drive(getattr(L, pre + "createDCtx"), getattr(L, pre + "freeDCtx"), getattr(L, pre + "decompressStream"),
                      lambda r: bool(getattr(L, pre + "isError")(r)), data, out_slice, in_chunk, dictionary, getattr(L, pre + "DCtx_loadDictionary"))

    def compare(data, out_slice, in_chunk=None, dictionary=None):
        ref = run(Z, "ZSTD_", data, out_slice, in_chunk, dictionary)
        got = run(K, "kzg_", data, out_slice, in_chunk, dictionary)
        assert got[1] == ref[1], (got[1], ref[1], len(got[0]), len(ref[0]))
        assert len(got[0]) == len(ref[0]) and got[0] == ref[0], (len(got[0]), len(ref[0]))
        if in_chunk is None:
            assert got[2] == ref[2]                 # the calls that returned 0: the frame boundaries
        return ref

    return compare


def test_frames_before_a_bad_frame_call_for_call(both, libzstd):
    """Four frames, the third with a flipped bit in its block: libzstd hands over the first two frames, then as much of the
    bad frame as whole calls deliver, then reports the error -- the same bytes and code here, for several slice sizes."""
    import corpus
    raws = [bytes(corpus.generate("logs", 600 + i, 30000)) for i in range(4)]
    for checksum in (True, False):
        frames = [bytearray(libzstd.compress(r, 3, checksum=checksum)) for r in raws]
        frames[2][len(frames[2]) // 2] ^= 0x10
        blob = b"".join(bytes(f) for f in frames)
        for out_slice in (8192, 100000, 30000, 1000):
            out, err, _ = both(blob, out_slice)              # (asserts: the same bytes and the same code as libzstd)
            assert out[:60000] == raws[0] + raws[1]
            if checksum:
                assert err == 22


def test_truncated_multiblock_frame_hands_over_its_complete_blocks(both, libzstd):
    """truncated input is not an error, and the output so far is exactly libzstd's -- the blocks that
    are complete (a streamed level-3 frame of five blocks, cut at many points; whole-input and pipe()-like pumps)."""
    import corpus
    raw = bytes(corpus.generate("mixed", 41, 600000))
    for pledged in (False, True):
        fr = libzstd.compress(raw, 3, checksum=True, pledged=pledged)
        cuts = [len(fr) * k // 23 for k in range(1, 23)] + [len(fr) - 1, len(fr) - 4, len(fr) - 5, 5, 9, 12]
        for cut in cuts:
            out, err, _ = both(fr[:cut], max(8192, cut // 10))
            assert err == 0 and raw.startswith(out)
        out, err, _ = both(fr[:len(fr) * 3 // 4], 8192, in_chunk=8192)
        assert err == 0 and raw.startswith(out) and len(out) >= 131072


def test_bad_block_in_the_middle_of_a_frame(both, libzstd):
    """A frame of several blocks with one broken block: the blocks in front of it are handed over first (as far as whole
    calls deliver them), then the error; a level-19 streamed frame with repeat-mode tables and treeless literals, too."""
    import corpus
    rng = np.random.default_rng(99)
    raw = bytes(corpus.generate("mixed", 43, 700000))
    for level, pledged in ((3, True), (19, False)):
        fr = libzstd.compress(raw, level, checksum=True, pledged=pledged)
        for _ in range(12):
            b = bytearray(fr)
            b[int(rng.integers(len(b) // 3, len(b) - 8))] ^= 1 << int(rng.integers(0, 8))
            for out_slice in (8192, 70000):
                both(bytes(b), out_slice)


def test_segmented_frames_roundtrip(api, both, libzstd):
    """Frames larger than one segment (32 MiB) and frames that arrive in pieces are decoded block run by block run with
    carried tables, repeat offsets, window and checksum: level 19 with an 8 MiB window, checksummed, through both pumps."""
    import corpus
    core, io, zstd = api
    raw = bytes(corpus.generate("mixed", 47, 6 << 20))
    fr = libzstd.compress(raw, 19, checksum=True, pledged=False)
    assert io.pipe(fr, zstd.ZstdDecompressor()).readByteArray() == raw
    out, err, zeros = both(fr, 65536, in_chunk=8192)
    assert err == 0 and out == raw and zeros[-1] == len(raw)
    big = bytes(corpus.generate("logs", 48, 40 << 20))               # > 32 MiB: segments even when it is all there
    frb = libzstd.compress(big, 3, checksum=True)
    assert core.transform(zstd.ZstdDecompressor(), frb) == big
    b = bytearray(frb); b[-2] ^= 0x40                                # checksum of a segmented frame
    with pytest.raises(zstd.IllegalStateException) as ei:
        core.transform(zstd.ZstdDecompressor(), bytes(b))
    assert ei.value.code == 22


def test_largeSample(api, libzstd):                      # ZstdTest.kt:67-82: 256 MiB + 3 random bytes, one frame
    """The one reference test round 1 did not replay.  The handle never holds more than a few segments of output."""
    import hashlib
    core, _, zstd = api
    n = 256 * 1024 * 1024 + 3
    rng = np.random.default_rng(2026)
    data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    compressed = libzstd.compress(data, 3)
    out = core.transform(zstd.ZstdDecompressor(), compressed)
    assert len(out) == n and hashlib.sha256(out).digest() == hashlib.sha256(data).digest()


def test_raw_blocks_arrive_in_pieces(both, libzstd):
    """What the stream fuzzer (tests/stress/gpu_fuzz_stream.py) found: libzstd passes raw blocks through as their bytes
    arrive, reports a block's input taken when it starts to hand the block over, and a pump that offers more input only
    when everything is taken sees both.  Incompressible frames (raw blocks), cut short or with a bad checksum, through
    pumps with several chunk and slice sizes."""
    rng = np.random.default_rng(373)
    small = rng.integers(0, 256, 50282, dtype=np.uint8).tobytes()
    big = rng.integers(0, 256, 700000, dtype=np.uint8).tobytes()
    f0 = libzstd.compress(small, 3, checksum=True)
    f1 = bytearray(libzstd.compress(big, 3, checksum=True))
    good = f0 + bytes(f1) + f0
    for out_slice, in_chunk in ((100000, 30000), (8192, 8192), (65536, 8192), (64, 65536), (1 << 20, 8192)):
        out, err, _ = both(good, out_slice, in_chunk=in_chunk)
        assert err == 0 and out == small + big + small
    f1[len(f1) // 3] ^= 0x04                             # inside a raw block: only the checksum notices
    bad = f0 + bytes(f1) + f0
    for out_slice, in_chunk in ((100000, 30000), (8192, 8192), (1000, None), (1 << 20, 8192), (64, 30000)):
        out, err, _ = both(bad, out_slice, in_chunk=in_chunk)
        assert err == 22 and out[:len(small)] == small
    for cut in (len(f0) + 10, len(f0) + 70000, len(f0) + 131080, len(good) - 2):
        for out_slice, in_chunk in ((8192, None), (65536, 8192), (100000, 30000)):
            out, err, _ = both(good[:cut], out_slice, in_chunk=in_chunk)
            assert err == 0 and (small + big + small).startswith(out)
