"""Host-side mirrors of the reference's contract (ByteArraySlice, SliceTransform.transform(ByteArray),
pipe()) checked with a fake XOR transform, like the reference's XorSliceTransformTest
(kompressor-kotlinx-io/src/jvmTest/kotlin/com/ensody/kompressor/kotlinx/io/XorSliceTransformTest.kt)."""
import os

import pytest

from kompressor_b200.mirror.core import ByteArraySlice, SliceTransform, transform
from kompressor_b200.mirror.io import Buffer, BytesSource, pipe


class XorSliceTransform(SliceTransform):
    def __init__(self):
        self.calls = []

    def transform(self, input, output, finish):  # noqa: A002
        n = min(input.remainingRead, output.remainingWrite)
        self.calls.append((input.remainingRead, output.remainingWrite, finish))
        for i in range(n):
            output.data[output.writeStart + i] = input.data[input.readStart + i] ^ 0x5A
        input.readStart += n
        output.writeStart += n
        output.insufficient = input.hasData


DATA = os.urandom(16 * 1024 + 3)
EXPECTED = bytes(b ^ 0x5A for b in DATA)


def test_byte_array_slice_cursors():
    s = ByteArraySlice(bytearray(b"abcdef"))
    assert (s.readStart, s.writeStart, s.writeLimit, s.remainingRead, s.remainingWrite) == (0, 6, 6, 6, 0)
    assert s.hasData and s.isFull
    e = ByteArraySlice(8)
    assert (e.writeStart, e.writeLimit, e.remainingWrite, e.hasData) == (0, 8, 8, False)
    s.readInto(e, size=4)
    assert bytes(e.data[:4]) == b"abcd" and s.readStart == 4 and e.writeStart == 4
    s.moveReadableToStart()
    assert (s.readStart, s.writeStart) == (0, 2) and bytes(s.data[:2]) == b"ef"
    with pytest.raises(ValueError):
        s.readInto(e, size=99)


def test_transform_byte_array_call_pattern():
    t = XorSliceTransform()
    assert transform(t, DATA) == EXPECTED
    # SliceTransform.kt:35: blockSize = max(8192, n / 10); whole input visible on the first call; finish = true
    assert t.calls[0] == (len(DATA), 8192, True)
    assert all(c[2] for c in t.calls)
    assert transform(XorSliceTransform(), b"") == b""


def test_pipe_feeds_8k_slices():
    t = XorSliceTransform()
    assert pipe(DATA, t).readByteArray() == EXPECTED
    assert max(c[0] for c in t.calls) <= 8192 and max(c[1] for c in t.calls) <= 8192


def test_pipe_small_reads():
    src = pipe(BytesSource(DATA), XorSliceTransform())
    out = Buffer()
    got = b""
    while True:
        n = src.readAtMostTo(out, 100)
        if n < 0:
            break
        got += out.readByteArray()
    assert got == EXPECTED


def test_pipe_into_sink():
    """SliceTransform.pipe(RawSink) (SliceTransformRawSink.kt:10-47): 8 KiB slices in, finish on close()."""
    from kompressor_b200.mirror.io import BufferSink, transferTo
    t = XorSliceTransform()
    output = BufferSink()
    sink = pipe(t, output)
    assert transferTo(DATA, sink) == len(DATA)
    sink.close()
    assert output.closed and output.readByteArray() == EXPECTED
    assert max(c[0] for c in t.calls) <= 8192 and t.calls[-1][2] is True      # the last call is the finishing one
    with pytest.raises(RuntimeError):
        sink.write(Buffer(b"x"), 1)
    # two transforms chained through sinks
    output = BufferSink()
    sink = pipe(XorSliceTransform(), pipe(XorSliceTransform(), output))
    transferTo(DATA, sink)
    sink.close()
    assert output.readByteArray() == DATA
