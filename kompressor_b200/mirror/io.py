"""Host-side mirror of kompressor-kotlinx-io's pipe() adaptors (8 KiB in, segment-sized out).

Reference:
  RawSource.pipe(SliceTransform)        kompressor-kotlinx-io/src/commonMain/kotlin/com/ensody/kompressor/kotlinx/io/SliceTransformRawSource.kt:12-55
  BufferSliceTransformHelper.transform  .../BufferSliceTransformHelper.kt:26-42
  transform(input, BufferTracker)       .../BufferExt.kt:41-59  (writeToTail with minCapacity 64, then 4096)
  SliceTransform.pipe(RawSink)          .../SliceTransformRawSink.kt:10-11
kotlinx-io's Buffer is modelled as a list of 8 KiB segments, which is what bounds every slice the
transform sees to 8192 bytes in and at most 8192 bytes out per native call.
"""
from __future__ import annotations

import io as _io

from .core import ByteArraySlice, SliceTransform

SEGMENT_SIZE = 8192


class Buffer:
    """Minimal kotlinx-io Buffer: FIFO of bytes held in 8 KiB segments."""

    def __init__(self, data: bytes = b""):
        self._segs: list[list] = []   # [bytearray(SEGMENT_SIZE), pos, limit]
        self.size = 0
        if data:
            self.write(data)

    def exhausted(self) -> bool:
        return self.size == 0

    def write(self, data) -> None:
        mv = memoryview(data)
        while len(mv):
            if not self._segs or self._segs[-1][2] == SEGMENT_SIZE:
                self._segs.append([bytearray(SEGMENT_SIZE), 0, 0])
            seg = self._segs[-1]
            n = min(len(mv), SEGMENT_SIZE - seg[2])
            seg[0][seg[2]:seg[2] + n] = mv[:n]
            seg[2] += n
            self.size += n
            mv = mv[n:]

    def head(self):
        """(bytes, start, endExclusive) of the head segment (UnsafeBufferOperations.readFromHead)."""
        seg = self._segs[0]
        return seg[0], seg[1], seg[2]

    def consume_head(self, n: int) -> None:
        seg = self._segs[0]
        seg[1] += n
        self.size -= n
        if seg[1] == seg[2]:
            self._segs.pop(0)

    def tail_for_write(self, minCapacity: int):
        """UnsafeBufferOperations.writeToTail: a tail segment with at least minCapacity free bytes."""
        if not self._segs or SEGMENT_SIZE - self._segs[-1][2] < minCapacity:
            self._segs.append([bytearray(SEGMENT_SIZE), 0, 0])
        seg = self._segs[-1]
        return seg, seg[2], SEGMENT_SIZE

    def commit_tail(self, seg, n: int) -> None:
        seg[2] += n
        self.size += n
        if seg[2] == seg[1] and len(self._segs) > 1 and seg is self._segs[-1] and n == 0 and seg[2] == 0:
            self._segs.pop()

    def readByteArray(self) -> bytes:
        out = b"".join(bytes(s[0][s[1]:s[2]]) for s in self._segs)
        self._segs.clear()
        self.size = 0
        return out

    def readAtMostTo(self, sink: "Buffer", byteCount: int) -> int:
        if self.size == 0:
            return -1
        n = 0
        while n < byteCount and self.size:
            data, s, e = self.head()
            k = min(e - s, byteCount - n)
            sink.write(data[s:s + k])
            self.consume_head(k)
            n += k
        return n


class BytesSource:
    """RawSource over bytes / a file-like object."""

    def __init__(self, data):
        self._f = _io.BytesIO(data) if isinstance(data, (bytes, bytearray, memoryview)) else data

    def readAtMostTo(self, sink: Buffer, byteCount: int) -> int:
        chunk = self._f.read(byteCount)
        if not chunk:
            return -1
        sink.write(chunk)
        return len(chunk)

    def close(self) -> None:
        pass


class BufferTracker:
    def __init__(self, buffer: Buffer):
        self.buffer = buffer
        self.processed = 0
        self.insufficient = False


def _transform_into_tracker(transform: SliceTransform, input: ByteArraySlice, sink: BufferTracker, finish: bool) -> None:  # noqa: A002
    """BufferExt.kt:41-59: try the smallest possible tail segment first, retry once with >= 4096 bytes."""
    outputInsufficient = False
    while True:
        minCapacity = 4096 if outputInsufficient else 64
        seg, outStart, outEnd = sink.buffer.tail_for_write(minCapacity)
        output = ByteArraySlice(seg[0], outStart, outStart, outEnd)
        transform.transform(input, output, finish)
        written = output.writeStart - outStart
        sink.buffer.commit_tail(seg, written)
        outputInsufficient = (not outputInsufficient) and output.insufficient
        sink.processed += written
        sink.insufficient = output.insufficient
        if not outputInsufficient:
            break


class BufferSliceTransformHelper:
    def __init__(self, transform: SliceTransform):
        self.transform_ = transform
        self.inputBuffer = ByteArraySlice(16 * 1024)

    def transform(self, source: Buffer, sink: BufferTracker, readLimit: int, finish: bool) -> int:
        if source.exhausted():
            inBytes, inStart, inEnd = bytearray(0), 0, 0
        else:
            inBytes, inStart, inEnd = source.head()
        maxReadSize = min(readLimit, inEnd - inStart)
        inputSegment = ByteArraySlice(inBytes, inStart, inStart + maxReadSize)
        input_ = inputSegment
        if self.inputBuffer.hasData:
            inputSegment.readInto(self.inputBuffer)
            input_ = self.inputBuffer
        _transform_into_tracker(self.transform_, input_, sink, finish)
        inputInsufficient = input_.insufficient
        self.inputBuffer.moveReadableToStart()
        if inputInsufficient:
            inputSegment.readInto(self.inputBuffer)
        consumed = inputSegment.readStart - inStart
        if consumed:
            source.consume_head(consumed)
        return consumed


class SliceTransformRawSource:
    """RawSource.pipe(transform) (SliceTransformRawSource.kt:15-60)."""

    def __init__(self, rawSource, transform: SliceTransform):
        self.source = rawSource
        self.helper = BufferSliceTransformHelper(transform)
        self.exhaustedSource = False
        self.exhaustedTransform = False
        self.localInput = Buffer()
        self.blockSize = 8192
        self.closed = False

    def readAtMostTo(self, sink: Buffer, byteCount: int) -> int:
        if self.closed:
            raise RuntimeError("Source is closed")
        if byteCount <= 0:
            raise ValueError("At least 1 byte must be read")
        tracker = BufferTracker(sink)
        while not self.exhaustedTransform and tracker.processed < byteCount:
            remaining = self.blockSize - self.localInput.size
            if not self.exhaustedSource and remaining > 1024 and self.source.readAtMostTo(self.localInput, remaining) < 0:
                self.exhaustedSource = True
            if not self.exhaustedTransform:
                self.helper.transform(self.localInput, tracker, byteCount - tracker.processed, self.exhaustedSource)
                if self.exhaustedSource and self.localInput.exhausted() and not tracker.insufficient:
                    self.exhaustedTransform = True
        return -1 if (tracker.processed == 0 and self.exhaustedTransform) else tracker.processed

    def readByteArray(self) -> bytes:
        """source.buffered().readByteArray()"""
        out = Buffer()
        while self.readAtMostTo(out, 8192) >= 0:
            pass
        return out.readByteArray()

    def close(self) -> None:
        self.source.close()
        self.closed = True


class BufferSink:
    """RawSink that collects what is written to it (kotlinx-io: a Buffer used as a Sink)."""

    def __init__(self):
        self.buffer = Buffer()
        self.closed = False

    def write(self, source: Buffer, byteCount: int) -> None:
        source.readAtMostTo(self.buffer, byteCount)

    def flush(self) -> None:
        pass

    def close(self) -> None:
        self.closed = True

    def readByteArray(self) -> bytes:
        return self.buffer.readByteArray()


class SliceTransformRawSink:
    """SliceTransform.pipe(RawSink): RawSink (SliceTransformRawSink.kt:10-47).  What the transform produces goes to the
    downstream sink segment by segment; close() finishes the transform (finishInto, BufferSliceTransformHelper.kt:49-55)."""

    def __init__(self, transform: SliceTransform, rawSink):
        self.sink = rawSink
        self.helper = BufferSliceTransformHelper(transform)
        self.closed = False

    def _drain(self, staged: Buffer) -> None:
        if staged.size:
            self.sink.write(staged, staged.size)

    def write(self, source: Buffer, byteCount: int) -> None:
        if self.closed:
            raise RuntimeError("Sink is closed")
        if byteCount <= 0:
            return
        bytesRead = 0
        while bytesRead < byteCount:
            staged = Buffer()                         # sink.writeToInternalBufferReturning { ... }
            bytesRead += self.helper.transform(source, BufferTracker(staged), byteCount - bytesRead, False)
            self._drain(staged)

    def flush(self) -> None:
        if self.closed:
            raise RuntimeError("Sink is closed")
        self.sink.flush()

    def close(self) -> None:
        if self.closed:
            return
        while True:                                   # finishInto: consume until the transform has nothing left
            staged = Buffer()
            tracker = BufferTracker(staged)
            _transform_into_tracker(self.helper.transform_, self.helper.inputBuffer, tracker, True)
            self._drain(staged)
            if not tracker.insufficient:
                break
        if self.helper.inputBuffer.insufficient:
            raise RuntimeError("Insufficient data in input buffer")
        if self.helper.inputBuffer.hasData:
            raise RuntimeError("Input buffer was not fully consumed after close()")
        self.closed = True
        self.sink.close()


def transferTo(data, sink) -> int:
    """Source.transferTo(sink) for bytes: everything is written through 8 KiB segments."""
    buf = Buffer(bytes(data))
    n = buf.size
    while buf.size:
        sink.write(buf, min(buf.size, SEGMENT_SIZE))
    return n


def pipe(source, transform):
    """RawSource.pipe(transform): RawSource  (source first)  /  SliceTransform.pipe(sink): RawSink  (transform first)."""
    if isinstance(source, SliceTransform):
        return SliceTransformRawSink(source, transform)
    if isinstance(source, (bytes, bytearray, memoryview)):
        source = BytesSource(source)
    return SliceTransformRawSource(source, transform)
