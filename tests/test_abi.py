"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/kzg.h
declares; host-only entry points behave (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kzg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kzg_[A-Za-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(kzg_lib):
    from kompressor_b200._lib import EXPORTS
    syms = declared_symbols()
    assert len(syms) >= 16
    assert sorted(EXPORTS) == syms
    for s in syms:
        assert hasattr(kzg_lib, s), s


def test_library_is_built_for_sm_100a():
    import subprocess
    from kompressor_b200._lib import LIB_PATH
    out = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_error_names_and_codes(kzg_lib, libzstd):
    for code in (1, 10, 12, 14, 16, 20, 22, 24, 30, 32, 40, 42, 44, 46, 62, 64, 66, 70, 72, 74, 80, 82):
        assert kzg_lib.kzg_getErrorName((1 << 64) - code).decode() == libzstd.error_name(-code), code
    assert kzg_lib.kzg_isError((1 << 64) - 20) == 1 and kzg_lib.kzg_isError(12345) == 0
    assert kzg_lib.kzg_isError((1 << 64) - 120) == 0      # libzstd: error iff > (size_t)-120


def test_scan_frames_matches_oracle(kzg_lib, oracle_mod, manifest):
    from helpers import gbytes
    from kompressor_b200.batch import scan_frames
    names = [v["name"] for v in manifest["vectors"] if v["name"] not in ("skippable_then_hello", "two_frames")]
    blob = b"".join(gbytes(n + ".zst") for n in names)
    off, size, content, consumed = scan_frames(blob)
    assert consumed == len(blob) and len(off) == len(names)
    pos = 0
    for i, n in enumerate(names):
        assert int(off[i]) == pos and int(size[i]) == oracle_mod.find_frame_size(blob[pos:]) == len(gbytes(n + ".zst"))
        pos += int(size[i])
    sizes = {v["name"]: v["size"] for v in manifest["vectors"]}
    for i, n in enumerate(names):
        assert int(content[i]) >= sizes[n]
    # truncated tail: stops at the last complete frame, no error
    off2, _, _, consumed2 = scan_frames(blob[:-3])
    assert len(off2) == len(names) - 1 and consumed2 == int(off[-1])
    # unknown magic -> error -10
    from kompressor_b200 import KzgError
    with pytest.raises(KzgError):
        scan_frames(b"garbage!garbage!")


def test_no_cpu_fallback_without_gpu(kzg_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    assert kzg_lib.kzg_ctx_create(ctypes.byref(h), -1) == 64 and not h
    assert not kzg_lib.kzg_createDCtx()
    from kompressor_b200.mirror.zstd import IllegalStateException, ZstdDecompressor
    with pytest.raises(IllegalStateException):
        ZstdDecompressor()


def test_product_never_touches_oracle():
    pkg = os.path.join(ROOT, "kompressor_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text, f
                assert "hostsim" not in text or f.endswith(".cuh"), f


def test_jni_shim_typechecks_and_names_the_reference_symbols():
    """Layer C (kzg_jni.cpp) cannot be linked without a JDK; check that it compiles against the minimal
    jni.h stub and defines exactly the symbols Kompressor's ZstdWrapper binds (all ten externals of libzstd-jni.so)
    (kompressor-zstd--nativelib/src/jvmCommonMain/kotlin/com/ensody/kompressor/zstd/ZstdWrapper.kt:24-60)."""
    import subprocess
    src = os.path.join(ROOT, "kompressor_b200", "csrc", "kzg_jni.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "tests", "jni_stub"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(src).read()
    defined = set(re.findall(r"JNICALL\s+(Java_com_ensody_kompressor_zstd_ZstdWrapper_\w+|JNI_On\w+)\s*\(", text))
    assert defined == {"JNI_OnLoad", "JNI_OnUnload"} | {"Java_com_ensody_kompressor_zstd_ZstdWrapper_" + n for n in (
        "createDecompressor", "freeDecompressor", "loadDecompressorDictionary", "decompressStream", "getErrorName",
        "createCompressor", "freeCompressor", "setParameter", "loadCompressorDictionary", "compressStream")}


def test_seek_table_roundtrip(kzg_lib, libzstd):
    """zstd seekable-format index (host code, no GPU): written by kzg_write_seek_table, used by kzg_scan_frames when it
    describes the buffer exactly, ignored (plain walk) when it does not; libzstd skips it like any skippable frame."""
    import ctypes
    import corpus
    from kompressor_b200._lib import lib
    from kompressor_b200.batch import scan_frames
    L = lib()
    raws = [bytes(corpus.generate("logs", 900 + i, 3000 + 777 * i)) for i in range(9)]
    frames = [libzstd.compress(r, 3, checksum=bool(i & 1)) for i, r in enumerate(raws)]
    body = b"".join(frames)
    n = len(frames)
    comp = np.array([len(f) for f in frames], np.uint64)
    cont = np.array([len(r) for r in raws], np.uint64)
    cap = int(L.kzg_seek_table_bound(n))
    assert cap == 8 + 8 * n + 9
    tab = np.zeros(cap, np.uint8)
    assert L.kzg_write_seek_table(tab.ctypes.data, cap, comp.ctypes.data, cont.ctypes.data, n) == cap
    assert L.kzg_isError(L.kzg_write_seek_table(tab.ctypes.data, cap - 1, comp.ctypes.data, cont.ctypes.data, n))
    table = tab.tobytes()
    assert table[:4] == bytes.fromhex("5e2a4d18") and table[-4:] == bytes.fromhex("b1ea928f")
    blob = body + table
    off, size, content, consumed = scan_frames(blob)
    assert consumed == len(blob) and len(off) == n + 1
    assert list(size[:n]) == list(comp) and list(content[:n]) == list(cont)
    assert int(off[n]) == len(body) and int(size[n]) == cap and int(content[n]) == 0
    # the same table in front of one more frame does not describe the buffer: plain walk, same frames
    off2, size2, content2, consumed2 = scan_frames(blob + frames[0])
    assert len(off2) == n + 2 and list(size2[:n]) == list(comp) and consumed2 == len(blob) + len(frames[0])
    # an entry that does not match: plain walk again
    bad = bytearray(blob)
    bad[len(body) + 8] ^= 1
    off3, size3, _, _ = scan_frames(bytes(bad))
    assert list(size3[:n]) == list(comp)
    out, err = libzstd.decompress_oneshot(blob, sum(len(r) for r in raws) + 64)
    assert err == 0 and out == b"".join(raws)


def test_seek_table_with_checksum_entries(kzg_lib, libzstd):
    """A seek table written by another tool may carry a checksum per entry (descriptor bit 7: 12-byte entries)."""
    import struct
    from kompressor_b200.batch import scan_frames
    raws = [bytes([65 + i]) * (1000 + i) for i in range(5)]
    frames = [libzstd.compress(r, 3) for r in raws]
    body = b"".join(frames)
    entries = b"".join(struct.pack("<III", len(f), len(r), 0xDEADBEEF) for f, r in zip(frames, raws))
    table = struct.pack("<II", 0x184D2A5E, len(entries) + 9) + entries + struct.pack("<IBI", len(frames), 0x80, 0x8F92EAB1)
    off, size, content, consumed = scan_frames(body + table)
    assert consumed == len(body) + len(table) and len(off) == 6
    assert [int(x) for x in content[:5]] == [len(r) for r in raws] and int(size[5]) == len(table)
    # reserved descriptor bits set: not a table this reader accepts -> plain walk, same result
    bad = body + table[:-5] + bytes([0x84]) + table[-4:]
    off2, size2, _, _ = scan_frames(bad)
    assert len(off2) == 6 and [int(x) for x in size2[:5]] == [len(f) for f in frames]


def test_scan_frames_bounds_a_hostile_content_size(kzg_lib, libzstd):
    """Frame_Content_Size is untrusted: a 17-byte frame that claims 2^64 - 2 bytes in front of a valid frame
    must not wrap any capacity arithmetic -- the capacity handed out is bounded by what the frame's blocks can
    regenerate, so the frame later fails like in libzstd (FCS larger than the decoded size -> corruption_detected)."""
    import struct
    from kompressor_b200.batch import scan_frames
    hostile = struct.pack("<IB", 0xFD2FB528, 0xE0) + struct.pack("<Q", 0xFFFFFFFFFFFFFFFE) + bytes([0x01, 0x00, 0x00])   # single segment, 8-byte FCS, empty last raw block
    good_raw = bytes(range(256)) * 40
    good = libzstd.compress(good_raw, 3)
    off, size, content, consumed = scan_frames(hostile + good)
    assert consumed == len(hostile) + len(good) and len(off) == 2
    assert int(size[0]) == len(hostile) and int(content[0]) == 0          # its only block regenerates nothing
    assert int(content[1]) == len(good_raw)
    out, err = libzstd.decompress_oneshot(hostile, 64)
    assert err == 20
    # an RLE block cannot claim more than the block maximum either (4 bytes of input, 2 MiB claimed)
    rle = struct.pack("<IB", 0xFD2FB528, 0x20) + bytes([0x00]) + struct.pack("<I", 1 | (1 << 1) | (0x1FFFFF << 3))[:3] + b"x"
    _, _, c3, _ = scan_frames(rle)
    assert int(c3[0]) <= 128 * 1024


def test_seek_table_is_only_a_hint(kzg_lib, libzstd):
    """A stale seek table (sizes that do not match the frames' own headers) must not change what is decoded: libzstd
    ignores skippable frames, so the block headers are walked as if the table were not there."""
    import struct
    from kompressor_b200.batch import scan_frames
    raws = [bytes([97 + i]) * (500 + 100 * i) + bytes(range(200)) for i in range(4)]
    frames = [libzstd.compress(r, 3) for r in raws]
    body = b"".join(frames)

    def table(entries):
        e = b"".join(struct.pack("<II", c, d) for c, d in entries)
        return struct.pack("<II", 0x184D2A5E, len(e) + 9) + e + struct.pack("<IBI", len(entries), 0, 0x8F92EAB1)

    honest = [(len(f), len(r)) for f, r in zip(frames, raws)]
    understated = [(c, d - 1 if i == 2 else d) for i, (c, d) in enumerate(honest)]
    merged = [honest[0], (honest[1][0] + honest[2][0], honest[1][1] + honest[2][1]), honest[3]]
    for entries in (understated, merged):
        off, size, content, consumed = scan_frames(body + table(entries))
        assert len(off) == 5 and consumed == len(body) + len(table(entries))
        assert [int(x) for x in size[:4]] == [len(f) for f in frames]
        assert [int(x) for x in content[:4]] == [len(r) for r in raws]
    off, size, content, _ = scan_frames(body + table(honest))
    assert [int(x) for x in content[:4]] == [len(r) for r in raws]
