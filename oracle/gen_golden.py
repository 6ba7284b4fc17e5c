"""TEST INFRASTRUCTURE: generates tests/golden/ with the real libzstd 1.5.7 (oracle/libzstd_ref.py).

Run in the build container:  python oracle/gen_golden.py
Every fixture records the frame bytes, the expected output (or its sha256 when large) and, for the
error vectors of SURVEY.md Appendix C, the error code libzstd returns through the reference's call
pattern (ZSTD_decompressStream driven like SliceTransform.transform(ByteArray)).
"""
from __future__ import annotations

import base64
import hashlib
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import corpus  # noqa: E402
from oracle.libzstd_ref import load  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main() -> None:
    z = load("1.5.7")
    assert z.version == "1.5.7", z.version
    os.makedirs(OUT, exist_ok=True)
    man = {"libzstd": z.version, "vectors": [], "errors": []}

    def add(name, frame, raw, note, dictionary=None, keep_raw=True):
        out, err, _ = z.decompress_stream(frame, dictionary=dictionary)
        assert err == 0 and out == raw, name
        with open(os.path.join(OUT, name + ".zst"), "wb") as f:
            f.write(frame)
        e = {"name": name, "note": note, "size": len(raw), "sha256": hashlib.sha256(raw).hexdigest()}
        if keep_raw:
            with open(os.path.join(OUT, name + ".raw"), "wb") as f:
                f.write(raw)
        if dictionary is not None:
            e["dictionary_b64"] = base64.b64encode(dictionary).decode()
        man["vectors"].append(e)

    # --- the reference's own known-answer vectors
    hello = base64.b64decode("KLUv/QRYuQAAaGVsbG8gY29tcHJlc3Npb24gd29ybGR8Qm9f")
    add("ref_sampleHello", hello, b"hello compression world", "ZstdTest.kt:84-91 (zstd CLI output, raw block + XXH64 trailer)")
    d = b"this is a test dictionary"
    data = b"this is some data to compress using a dictionary"
    fr = z.compress(data, 3, dictionary=d)
    assert fr.hex() == "28b52ffd2030350100e8736f6d65206461746120746f20636f6d7072657373207573696e67206102004d659c890904"
    add("ref_dictionary", fr, data, "ZstdTest.kt:49-65 dictionary pair, raw-content dictionary", dictionary=d)
    add("empty", z.compress(b"", 3), b"", "empty input -> 28b52ffd2000010000")
    # --- stage coverage
    add("logs_8k_l3", z.compress(bytes(corpus.generate("logs", 1, 8192)), 3), bytes(corpus.generate("logs", 1, 8192)),
        "Huffman x4 literals + three FSE tables")
    add("json_8k_l3_sum", z.compress(bytes(corpus.generate("json", 2, 8192)), 3, checksum=True),
        bytes(corpus.generate("json", 2, 8192)), "as above with content checksum")
    small = bytes(corpus.generate("logs", 3, 700))
    add("logs_700_l3", z.compress(small, 3), small, "single-stream Huffman / predefined tables")
    zeros = bytes(40000)
    add("zeros_40000", z.compress(zeros, 3), zeros, "RLE literals + long overlapping match")
    lit = bytes(corpus.generate("literal", 5, 65536))
    add("literal_64k_l1", z.compress(lit, 1), lit, "literal-heavy level 1")
    rnd = bytes(corpus.generate("random", 6, 131075))
    add("random_131075_l3", z.compress(rnd, 3), rnd, "README round trip shape: raw blocks 131072 + 3", keep_raw=False)
    rnd_s = z.compress(rnd[:40000], 3, pledged=False)
    add("random_40000_streamed", rnd_s, rnd[:40000], "pipe() shape: window descriptor, no FCS, raw blocks", keep_raw=False)
    mixed = bytes(corpus.generate("mixed", 7, 400000))
    add("mixed_400k_l19_streamed", z.compress(mixed, 19, checksum=True, pledged=False), mixed,
        "multi-block, treeless literals, repeat-mode tables, cross-block matches", keep_raw=False)
    add("mixed_400k_l3", z.compress(mixed, 3), mixed, "4 blocks at level 3", keep_raw=False)
    # formatted dictionary (ZDICT_trainFromBuffer): entropy tables + repeat offsets + content, dictID in the frame
    fdict = z.train_dictionary([bytes(corpus.generate("json", 100 + i, 2000)) for i in range(200)], 4096)
    assert fdict[:4] == bytes.fromhex("37a430ec")
    djson = bytes(corpus.generate("json", 999, 3000))
    add("dict_formatted_json_3k", z.compress(djson, 3, dictionary=fdict), djson,
        "formatted dictionary: treeless literals / repeat-mode tables / repeat offsets come from the dictionary", dictionary=fdict)
    # hand-made: pure RLE block frame and skippable + frame
    rle = bytes.fromhex("28b52ffd") + bytes([0x20, 100]) + struct.pack("<I", (100 << 3) | (1 << 1) | 1)[:3] + b"Z"
    add("rle_block_100", rle, b"Z" * 100, "single RLE block (hand-made)")
    skip = struct.pack("<II", 0x184D2A53, 5) + b"abcde"
    add("skippable_then_hello", skip + hello, b"hello compression world", "skippable frame is ignored")
    add("two_frames", hello + z.compress(small, 3), b"hello compression world" + small, "two concatenated frames")

    # --- error vectors (Appendix C): expected = what libzstd returns through the reference call pattern
    def add_err(name, frame, note):
        out, err, _ = z.decompress_stream(frame)
        with open(os.path.join(OUT, name + ".zst"), "wb") as f:
            f.write(frame)
        man["errors"].append({"name": name, "note": note, "error": err, "out_size": len(out),
                              "out_sha256": hashlib.sha256(out).hexdigest()})

    js = z.compress(bytes(corpus.generate("json", 2, 8192)), 3, checksum=True)
    b = bytearray(js); b[-1] ^= 0xFF
    add_err("err_checksum_flipped", bytes(b), "checksum byte flipped -> -22")
    b = bytearray(js); b[0] ^= 0xFF
    add_err("err_magic", bytes(b), "magic corrupted -> -10")
    b = bytearray(js); b[len(b) // 2] ^= 0x55
    add_err("err_corrupt_block", bytes(b), "byte flipped in a compressed block")
    add_err("err_truncated", js[:len(js) // 2], "truncated in the middle: no error, no output")
    add_err("err_missing_checksum_byte", js[:-1], "last checksum byte missing: no error (checksum never verified)")
    add_err("err_trailing_garbage", hello + b"garbage!", "frame then trailing garbage -> -10")
    add_err("err_reserved_block", bytes.fromhex("28b52ffd") + bytes([0x20, 10]) + struct.pack("<I", (10 << 3) | (3 << 1) | 1)[:3] + b"0123456789", "reserved block type 3 -> -20")
    fcs_bad = bytearray(z.compress(small, 3)); fcs_bad[5] ^= 0x01
    add_err("err_fcs_mismatch", bytes(fcs_bad), "FCS differs from decoded size -> -20")
    add_err("err_windowlog_28", bytes.fromhex("28b52ffd") + bytes([0x00, (28 - 10) << 3]) + bytes([0x01, 0x00, 0x00]), "windowLog 28 -> -16")
    add_err("err_needs_dictionary", fr, "frame needs the dictionary, none loaded -> -20")
    add_err("err_reserved_bit", bytes.fromhex("28b52ffd") + bytes([0x28, 1]) + bytes([0x09, 0x00, 0x00, 0x41]), "reserved header bit -> -14")

    # --- Huffman streams that are not consumed exactly.  Single-bit flips inside the literal streams of
    # one frame, found by a seeded search; libzstd's verdict is recorded, and for the one documented divergence also ours.
    import numpy as np
    import oracle
    man["divergences"] = []
    hraw = bytes(corpus.generate("json", 11, 20000))
    rng = np.random.default_rng(2026)
    want = {"huf_leftover_bits": None, "err_huf_ran_out_checksum": None, "div_huf_ran_out_no_checksum": None}
    for checksum in (False, True):
        hfr = z.compress(hraw, 3, checksum=checksum)
        for _ in range(4000):
            bb = bytearray(hfr)
            pos = int(rng.integers(40, len(bb) // 3))          # the literals section comes first in the block
            bb[pos] ^= 1 << int(rng.integers(0, 8))
            ref, rerr = z.decompress_oneshot(bytes(bb), 65536)
            out, err = oracle.decompress(bytes(bb), 65536)
            if not checksum and rerr == 0 and err == 0 and ref != hraw and want["huf_leftover_bits"] is None:
                want["huf_leftover_bits"] = (bytes(bb), ref)
            if not checksum and rerr == 0 and err == 20 and want["div_huf_ran_out_no_checksum"] is None:
                want["div_huf_ran_out_no_checksum"] = (bytes(bb), ref)
            if checksum and rerr == 22 and err == 22 and want["err_huf_ran_out_checksum"] is None:
                # make sure it is the literals that differ, not the trailer: the same flip without the verdict rule is -20
                want["err_huf_ran_out_checksum"] = (bytes(bb), None)
            if all(v is not None for v in want.values()):
                break
    assert all(v is not None for v in want.values()), want
    fr_l, ref_l = want["huf_leftover_bits"]
    assert oracle.decompress(fr_l, 65536) == (ref_l, 0)
    add("huf_leftover_bits", fr_l, ref_l, "a flipped bit leaves bits over in a Huffman stream: libzstd's fast loops accept, same bytes", keep_raw=False)
    add_err("err_huf_ran_out_checksum", want["err_huf_ran_out_checksum"][0],
            "a Huffman stream runs out of bits, frame with checksum: libzstd makes up literals and reports -22")
    fr_d, ref_d = want["div_huf_ran_out_no_checksum"]
    with open(os.path.join(OUT, "div_huf_ran_out_no_checksum.zst"), "wb") as f:
        f.write(fr_d)
    man["divergences"].append({"name": "div_huf_ran_out_no_checksum", "libzstd_error": 0, "libzstd_out_sha256": hashlib.sha256(ref_d).hexdigest(),
                               "ours": 20, "note": "a Huffman stream runs out of bits, no checksum: libzstd accepts made-up literals, "
                                                   "the oracle and the GPU path report corruption_detected (documented)"})

    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(man, f, indent=1)
    tot = sum(os.path.getsize(os.path.join(OUT, x)) for x in os.listdir(OUT))
    print(f"wrote {len(man['vectors'])} vectors, {len(man['errors'])} error vectors, {tot/1024:.0f} KiB")
    for e in man["errors"]:
        print(" ", e["name"], e["error"], e["out_size"])


if __name__ == "__main__":
    main()
