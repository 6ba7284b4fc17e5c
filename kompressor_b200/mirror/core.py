"""Host-side mirror of Kompressor's core contract (kompressor-core), same names and semantics.

Reference:
  ByteArraySlice     kompressor-core/src/commonMain/kotlin/com/ensody/kompressor/core/ByteArraySlice.kt:14-90
  SliceTransform     kompressor-core/src/commonMain/kotlin/com/ensody/kompressor/core/SliceTransform.kt:8-30
  transform(ByteArray), getOutput      SliceTransform.kt:33-56
Kotlin camelCase names are kept on purpose so the parity tests read like the reference's tests.
"""
from __future__ import annotations


class ByteArraySlice:
    """A slice of a byte array. Readable region [readStart, writeStart), writable [writeStart, writeLimit)."""

    __slots__ = ("data", "readStart", "writeStart", "writeLimit", "insufficient")

    def __init__(self, data, readStart: int = 0, writeStart: int | None = None, writeLimit: int | None = None,
                 insufficient: bool = False):
        if isinstance(data, int):          # ByteArraySlice(size): empty slice ready to be filled
            data = bytearray(data)
            writeStart = 0 if writeStart is None else writeStart
        elif not isinstance(data, bytearray):
            data = bytearray(data)
        self.data = data
        self.readStart = readStart
        self.writeStart = len(data) if writeStart is None else writeStart
        self.writeLimit = len(data) if writeLimit is None else writeLimit
        self.insufficient = insufficient

    @property
    def remainingRead(self) -> int:
        return self.writeStart - self.readStart

    @property
    def remainingWrite(self) -> int:
        return self.writeLimit - self.writeStart

    @property
    def hasData(self) -> bool:
        return self.remainingRead != 0

    @property
    def isFull(self) -> bool:
        return self.writeStart == self.writeLimit

    @property
    def shouldMoveReadable(self) -> bool:
        return (not self.hasData) or self.insufficient

    @property
    def shouldFlushWritten(self) -> bool:
        return self.isFull or self.insufficient

    def onInputConsumed(self) -> None:
        if self.shouldMoveReadable:
            self.moveReadableToStart()

    def moveReadableToStart(self) -> None:
        if self.hasData:
            n = self.remainingRead
            self.data[0:n] = self.data[self.readStart:self.writeStart]
        self.writeStart = self.remainingRead
        self.readStart = 0
        self.insufficient = False

    def setNewData(self, data=None, readStart: int = 0, writeStart: int = 0, writeLimit: int | None = None) -> None:
        self.data = bytearray(len(self.data)) if data is None else data
        self.readStart = readStart
        self.writeStart = writeStart
        self.writeLimit = len(self.data) if writeLimit is None else writeLimit
        self.insufficient = False

    def readInto(self, other, destOffset: int | None = None, size: int | None = None) -> None:
        """readInto(other: ByteArraySlice, size) or readInto(dest: bytearray, destOffset, size)."""
        if isinstance(other, ByteArraySlice):
            size = self.remainingRead if size is None else size
            if size < 0:
                raise ValueError("size must not be negative")
            if size > self.remainingRead:
                raise ValueError("size is larger than input slice")
            if size > other.remainingWrite:
                raise ValueError("size is larger than output slice")
            end = self.readStart + size
            other.data[other.writeStart:other.writeStart + size] = self.data[self.readStart:end]
            self.readStart = end
            other.writeStart += size
        else:
            size = self.remainingRead if size is None else size
            if size < 0:
                raise ValueError("size must not be negative")
            if size > self.remainingRead:
                raise ValueError("size is larger than input slice")
            if destOffset + size > len(other):
                raise ValueError("size + destOffset is larger than output size")
            end = self.readStart + size
            other[destOffset:destOffset + size] = self.data[self.readStart:end]
            self.readStart = end

    def __repr__(self) -> str:
        return f"ByteArraySlice(remainingRead={self.remainingRead}, insufficient={self.insufficient})"


class SliceTransform:
    """Transforms data contained in ByteArraySlices (SliceTransform.kt:8-30)."""

    def transform(self, input: ByteArraySlice, output: ByteArraySlice, finish: bool) -> None:  # noqa: A002
        raise NotImplementedError


def getOutput(slices) -> bytes:
    """List<ByteArraySlice>.getOutput (SliceTransform.kt:47-56)."""
    return b"".join(bytes(s.data[s.readStart:s.writeStart]) for s in slices)


def transform(t: SliceTransform, input) -> bytes:  # noqa: A002
    """SliceTransform.transform(input: ByteArray): ByteArray (SliceTransform.kt:33-45), same loop:
    whole input visible on the first call, output in max(8192, n/10) slices, finish = true."""
    inputSlice = ByteArraySlice(input if isinstance(input, bytearray) else bytearray(input))
    blockSize = max(8192, len(inputSlice.data) // 10)
    outputSlices = [ByteArraySlice(blockSize)]
    while True:
        outputSlice = outputSlices[-1]
        t.transform(inputSlice, outputSlice, True)
        if outputSlice.insufficient:
            outputSlices.append(ByteArraySlice(blockSize))
        if not (inputSlice.remainingRead != 0 or outputSlice.insufficient):
            break
    return getOutput(outputSlices)
