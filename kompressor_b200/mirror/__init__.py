"""Python transliterations of the reference's Kotlin callers of the path -- `SliceTransform` / `ByteArraySlice`
(kompressor-core/.../SliceTransform.kt:8-56, ByteArraySlice.kt:14-90), the kotlinx-io `pipe()` pump
(kompressor-kotlinx-io/.../SliceTransformRawSource.kt:32-55, BufferExt.kt:41-59) and `ZstdDecompressor` / `ZstdCompressor`
(ZstdDecompressor.jvm.kt:10-40, ZstdCompressor.jvm.kt:10-51) -- bound to libkzg.so through ctypes.

They exist because this image has no JVM: the parity tests drive the C ABI through the same loops, slice sizes and error
handling the Kotlin code uses, so that they read like the reference's own tests.  They are test harness and usage
examples, not the deliverable: the product is the CUDA library behind include/kzg.h (see INTEGRATION.md for the binding a
Kompressor maintainer adds)."""
