// Zstandard header parsing: frame header, block header walk, literals-section header, Huffman tree
// description, FSE table descriptions, sequences-section header.  Serial per-thread code (KZG_HD).
// Format: RFC 8878 as restated in SURVEY.md Appendix A; accept/reject decisions follow libzstd 1.5.7,
// the library behind the reference's ZSTD_decompressStream call (Wrapper.cpp:178).
#pragma once
#include "kzg_bits.cuh"

struct KzgFrameHeader {
    u64 window_size, fcs;
    u32 hdr_size, dict_id, has_checksum, single_segment;
};

// Returns 0 or an error code. KZG_E_SRCSIZE_WRONG = header truncated.
KZG_HD u32 kzg_parse_frame_header(const u8* src, u64 size, KzgFrameHeader& h) {
    if (size < 5) return KZG_E_SRCSIZE_WRONG;
    if (kzg_rd32(src) != KZG_MAGIC) return KZG_E_PREFIX_UNKNOWN;
    const u32 fhd = src[4];
    const u32 fcs_flag = fhd >> 6, did_flag = fhd & 3;
    h.single_segment = (fhd >> 5) & 1;
    h.has_checksum = (fhd >> 2) & 1;
    if (fhd & 8) return KZG_E_FRAMEPARAM_UNSUPPORTED;
    const u32 did_bytes = did_flag == 3 ? 4 : did_flag;
    u32 fcs_bytes = fcs_flag == 0 ? 0 : (1u << fcs_flag);
    if (fcs_flag == 0 && h.single_segment) fcs_bytes = 1;
    h.hdr_size = 5 + (h.single_segment ? 0 : 1) + did_bytes + fcs_bytes;
    if (size < h.hdr_size) return KZG_E_SRCSIZE_WRONG;
    u32 pos = 5;
    h.window_size = 0;
    if (!h.single_segment) {
        const u32 b = src[pos++], wl = 10 + (b >> 3);
        if (wl > 31) return KZG_E_WINDOW_TOO_LARGE;
        h.window_size = (1ull << wl) + ((1ull << wl) >> 3) * (b & 7);
    }
    h.dict_id = 0;
    if (did_bytes == 1) h.dict_id = src[pos];
    else if (did_bytes == 2) h.dict_id = kzg_rd16(src + pos);
    else if (did_bytes == 4) h.dict_id = kzg_rd32(src + pos);
    pos += did_bytes;
    h.fcs = ~0ull;
    if (fcs_bytes == 1) h.fcs = src[pos];
    else if (fcs_bytes == 2) h.fcs = kzg_rd16(src + pos) + 256;
    else if (fcs_bytes == 4) h.fcs = kzg_rd32(src + pos);
    else if (fcs_bytes == 8) h.fcs = kzg_rd64(src + pos);
    if (h.single_segment) h.window_size = h.fcs;
    return KZG_OK;
}

// ---------------------------------------------------------------------------------------------
// FSE normalised counts (Appendix A.5).  Returns bytes consumed, or 0 on error.
// Forward bit reader over a 64-bit window that is refilled a few bytes at a time (the description is read in
// fields of at most 10 bits: one load per field would be five times as many loads).
struct KzgFwdWin {
    const u8* p; u32 size;
    u64 win; u32 base;                      // win = bytes [base, base + 8) of p, bytes past `size` read as zero
    KZG_HD void fill(u32 byte) {
        base = byte;
        win = 0;
        for (u32 k = 0; k < 8; k++)
            if (byte + k < size) win |= (u64)p[byte + k] << (8 * k);
    }
    KZG_HD u32 bits(u32 bitpos, u32 n) {    // n <= 24
        const u32 byte = bitpos >> 3;
        if (byte < base || byte + 4 > base + 8) fill(byte);
        return (u32)(win >> (bitpos - 8 * base)) & ((1u << n) - 1u);
    }
};

KZG_HD u32 kzg_read_ncount(const u8* src, u32 size, u32 max_sym, u32 max_log, i16* norm, u32& nsym, u32& log) {
    if (size < 1) return 0;
    KzgFwdWin r;
    r.p = src; r.size = size;
    r.fill(0);
    u32 bitpos = 4;
    log = 5 + (src[0] & 15);
    if (log > max_log) return 0;
    i32 remaining = 1 << log;
    u32 sym = 0;
    while (remaining > 0) {
        if (sym > max_sym) return 0;
        const u32 bits = (u32)kzg_highbit((u32)remaining + 1) + 1;
        i32 v = (i32)r.bits(bitpos, bits);
        bitpos += bits;
        const i32 lower = (1 << (bits - 1)) - 1;
        const i32 thr = (1 << bits) - 1 - (remaining + 1);
        if ((v & lower) < thr) { bitpos -= 1; v &= lower; }
        else if (v > lower) v -= thr;
        const i32 prob = v - 1;
        remaining -= prob < 0 ? 1 : prob;
        if (remaining < 0) return 0;
        norm[sym++] = (i16)prob;
        if (prob == 0) {
            for (;;) {
                const u32 rep = r.bits(bitpos, 2);
                bitpos += 2;
                for (u32 i = 0; i < rep; i++) {
                    if (sym > max_sym) return 0;
                    norm[sym++] = 0;
                }
                if (rep != 3) break;
                if (bitpos > 8u * size + 64u) return 0;
            }
        }
    }
    const u32 used = (bitpos + 7) >> 3;
    if (used > size) return 0;
    nsym = sym;
    return used;
}

// Small FSE decode table used only for the FSE-coded Huffman weights (log <= 6).
struct KzgFseSmall { u8 sym[64]; u8 nb[64]; u8 base[64]; };

KZG_HD void kzg_fse_build_small(KzgFseSmall& t, const i16* norm, u32 nsym, u32 log) {
    const u32 size = 1u << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    u32 high = size, pos = 0;
    u8 next[256];
    for (u32 s = 0; s < nsym; s++) {
        if (norm[s] == -1) { t.sym[--high] = (u8)s; next[s] = 1; }
        else next[s] = (u8)norm[s];
    }
    for (u32 s = 0; s < nsym; s++)
        for (i32 i = 0; i < norm[s]; i++) {
            t.sym[pos] = (u8)s;
            do { pos = (pos + step) & mask; } while (pos >= high);
        }
    for (u32 i = 0; i < size; i++) {
        const u32 n = next[t.sym[i]]++;
        const u32 nb = log - (u32)kzg_highbit(n);
        t.nb[i] = (u8)nb;
        t.base[i] = (u8)((n << nb) - size);
    }
}

// Huffman tree description (Appendix A.4) -> weights incl. the implicit last one.
// Returns bytes consumed (>0) or 0 on error.
KZG_HD u32 kzg_read_huf_weights(const u8* src, u32 size, u8* w, u32& nsym_out, u32& maxbits_out) {
    if (size < 1) return 0;
    u32 n = 0, used;
    const u32 hb = src[0];
    if (hb >= 128) {
        n = hb - 127;
        used = 1 + (n + 1) / 2;
        if (used > size) return 0;
        for (u32 k = 0; 2 * k < n; k++) {               // two weights per byte, high nibble first
            const u32 b = src[1 + k];
            w[2 * k] = (u8)(b >> 4);
            if (2 * k + 1 < n) w[2 * k + 1] = (u8)(b & 15);
        }
    } else {
        used = 1 + hb;
        if (used > size) return 0;
        i16 norm[256];
        u32 nsym, log;
        const u32 hsz = kzg_read_ncount(src + 1, hb, 255, 6, norm, nsym, log);
        if (hsz == 0) return 0;
        KzgFseSmall t;
        kzg_fse_build_small(t, norm, nsym, log);
        KzgBwd b;
        if (!b.init(src + 1 + hsz, hb - hsz)) return 0;
        b.ensure32();
        u32 s1 = b.read(log), s2 = b.read(log);
        if (b.left < 0) return 0;
        for (;;) {
            if (n > 253) return 0;
            b.ensure32();
            w[n++] = t.sym[s1];
            s1 = t.base[s1] + b.read(t.nb[s1]);
            if (b.left < 0) { w[n++] = t.sym[s2]; break; }
            if (n > 253) return 0;
            w[n++] = t.sym[s2];
            s2 = t.base[s2] + b.read(t.nb[s2]);
            if (b.left < 0) { w[n++] = t.sym[s1]; break; }
        }
    }
    u32 total = 0, rank1 = 0;
    for (u32 i = 0; i < n; i++) {
        if (w[i] > KZG_HUF_MAXBITS + 1) return 0;
        if (w[i]) total += 1u << (w[i] - 1);
        rank1 += (w[i] == 1);
    }
    if (total == 0) return 0;
    const u32 maxbits = (u32)kzg_highbit(total) + 1;
    if (maxbits > KZG_HUF_MAXBITS) return 0;
    const u32 rest = (1u << maxbits) - total;
    if (rest & (rest - 1)) return 0;
    const u32 lastw = (u32)kzg_highbit(rest) + 1;
    w[n++] = (u8)lastw;
    rank1 += (lastw == 1);
    if (rank1 < 2 || (rank1 & 1)) return 0;
    nsym_out = n;
    maxbits_out = maxbits;
    return used;
}

// ---------------------------------------------------------------------------------------------
// Parses one compressed block's literals-section header and sequences-section header and the
// entropy descriptions between them; fills `cb` (everything except owners / pool offsets), `fd`, `hd`.
// src = block content, size = Block_Size.
KZG_HD void kzg_parse_cblock(const u8* src, u32 size, u32 block_max, KzgCBlock& cb, KzgFseDesc& fd, KzgHufDesc& hd) {
    cb.status_lit = cb.status_seq = cb.status_exec = KZG_OK;
    cb.lit_type = 0; cb.lit_regen = 0; cb.lit_src = 0; cb.n_streams = 0;
    cb.stream_size[0] = cb.stream_size[1] = cb.stream_size[2] = cb.stream_size[3] = 0;
    cb.n_seq = 0; cb.modes = 0; cb.seq_src = 0; cb.seq_size = 0; cb.regen = 0; cb.lit_used = 0;
    fd.kind[0] = fd.kind[1] = fd.kind[2] = 0;
    hd.nsym = 0; hd.maxbits = 0;
    u32 lit_total;   // bytes of the literals section
    // ---- literals section (A.3)
    if (size < 2) { cb.status_lit = KZG_E_CORRUPTION; return; }
    {
        const u32 type = src[0] & 3, fmt = (src[0] >> 2) & 3;
        cb.lit_type = type;
        if (type < 2) {
            u32 hs, regen;
            if (fmt == 0 || fmt == 2) { hs = 1; regen = src[0] >> 3; }
            else if (fmt == 1) { hs = 2; regen = kzg_rd16(src) >> 4; }
            else { if (size < 3) { cb.status_lit = KZG_E_CORRUPTION; return; } hs = 3; regen = kzg_rd24(src) >> 4; }
            if (regen > block_max) { cb.status_lit = KZG_E_CORRUPTION; return; }
            cb.lit_regen = regen; cb.lit_src = hs;
            if (type == 0) {
                if (hs + regen > size) { cb.status_lit = KZG_E_CORRUPTION; return; }
                lit_total = hs + regen;
            } else {
                if (fmt == 3 && size < 4) { cb.status_lit = KZG_E_CORRUPTION; return; }
                if (hs + 1 > size) { cb.status_lit = KZG_E_CORRUPTION; return; }
                lit_total = hs + 1;
            }
        } else {
            u32 hs, regen, csize, streams;
            if (size < 5) { cb.status_lit = KZG_E_CORRUPTION; return; }
            if (fmt == 0) { streams = 1; hs = 3; regen = (kzg_rd24(src) >> 4) & 0x3FF; csize = (kzg_rd24(src) >> 14) & 0x3FF; }
            else if (fmt == 1) { streams = 4; hs = 3; regen = (kzg_rd24(src) >> 4) & 0x3FF; csize = (kzg_rd24(src) >> 14) & 0x3FF; }
            else if (fmt == 2) { streams = 4; hs = 4; regen = (kzg_rd32(src) >> 4) & 0x3FFF; csize = kzg_rd32(src) >> 18; }
            else { streams = 4; hs = 5; regen = (kzg_rd32(src) >> 4) & 0x3FFFF; csize = (kzg_rd32(src) >> 22) | ((u32)src[4] << 10); }
            cb.n_streams = streams; cb.lit_regen = regen;
            if (regen > block_max) { cb.status_lit = KZG_E_CORRUPTION; return; }
            if (streams == 4 && regen < 6) { cb.status_lit = KZG_E_LITERALS_HEADER_WRONG; return; }
            if (hs + csize > size) { cb.status_lit = KZG_E_CORRUPTION; return; }
            u32 tsz = 0;
            if (type == 2) {
                tsz = kzg_read_huf_weights(src + hs, csize, hd.weight, hd.nsym, hd.maxbits);
                if (tsz == 0) { cb.status_lit = KZG_E_CORRUPTION; return; }
            }
            const u32 rem = csize - tsz;
            const u8* p = src + hs + tsz;
            if (streams == 1) {
                cb.lit_src = hs + tsz;
                cb.stream_size[0] = rem;
                if (rem == 0) { cb.status_lit = KZG_E_CORRUPTION; return; }
            } else {
                if (rem < 10) { cb.status_lit = KZG_E_CORRUPTION; return; }
                const u32 l1 = kzg_rd16(p), l2 = kzg_rd16(p + 2), l3 = kzg_rd16(p + 4);
                if (l1 + l2 + l3 + 6 > rem) { cb.status_lit = KZG_E_CORRUPTION; return; }
                const u32 seg = (regen + 3) / 4;
                if (3 * seg > regen) { cb.status_lit = KZG_E_CORRUPTION; return; }
                cb.lit_src = hs + tsz + 6;
                cb.stream_size[0] = l1; cb.stream_size[1] = l2; cb.stream_size[2] = l3;
                cb.stream_size[3] = rem - 6 - l1 - l2 - l3;
                if (l1 == 0 || l2 == 0 || l3 == 0 || cb.stream_size[3] == 0) { cb.status_lit = KZG_E_CORRUPTION; return; }
            }
            lit_total = hs + csize;
        }
    }
    // ---- sequences section header (A.6)
    const u8* p = src + lit_total;
    const u8* const end = src + size;
    if (p >= end) { cb.status_seq = KZG_E_SRCSIZE_WRONG; return; }
    u32 nseq = *p++;
    if (nseq > 0x7F) {
        if (nseq == 0xFF) { if (p + 2 > end) { cb.status_seq = KZG_E_SRCSIZE_WRONG; return; } nseq = kzg_rd16(p) + 0x7F00; p += 2; }
        else { if (p >= end) { cb.status_seq = KZG_E_SRCSIZE_WRONG; return; } nseq = ((nseq - 0x80) << 8) + *p++; }
    }
    cb.n_seq = nseq;
    if (nseq == 0) { if (p != end) cb.status_seq = KZG_E_CORRUPTION; return; }
    if (p + 1 > end) { cb.status_seq = KZG_E_SRCSIZE_WRONG; return; }
    const u32 mb = *p++;
    if (mb & 3) { cb.status_seq = KZG_E_CORRUPTION; return; }
    const u32 m[3] = { mb >> 6, (mb >> 4) & 3, (mb >> 2) & 3 };   // LL, OF, ML
    cb.modes = m[0] | (m[1] << 2) | (m[2] << 4);
    const u32 max_sym[3] = { 35, 31, 52 }, max_log[3] = { 9, 8, 9 };
    for (u32 t = 0; t < 3; t++) {
        if (m[t] == 1) {
            if (p >= end || *p > max_sym[t]) { cb.status_seq = KZG_E_CORRUPTION; return; }
            fd.kind[t] = 1; fd.log[t] = 0; fd.nsym[t] = 1; fd.norm[t][0] = (i16)*p++;
        } else if (m[t] == 2) {
            u32 nsym, log;
            const u32 h = kzg_read_ncount(p, (u32)(end - p), max_sym[t], max_log[t], fd.norm[t], nsym, log);
            if (h == 0) { cb.status_seq = KZG_E_CORRUPTION; return; }
            fd.kind[t] = 2; fd.log[t] = (u8)log; fd.nsym[t] = (u8)nsym;
            p += h;
        }
    }
    cb.seq_src = (u32)(p - src);
    cb.seq_size = (u32)(end - p);
    if (cb.seq_size == 0 || end[-1] == 0) cb.status_seq = KZG_E_CORRUPTION;
}
