// Batch frame writer for sm_100a ("the step before the path"): compresses many
// independent chunks (<= 128 KiB each, 64 KiB by default) into one single-block Zstandard frame each, so that
// producers emit the batch shape the decoder is built for.  What the reference does here is
// ZSTD_compressStream2 (kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:75-121); any conforming
// encoder is interchangeable with it, because decoders only see the format (RFC 8878).
// Parity bar for this side: libzstd 1.5.7 decodes every emitted frame to the original bytes.
//
// Format choices of this version: one compressed block per frame, Huffman-coded literals (4 streams), sequence tables made for the block (FSE mode; RLE mode for a
// single code; the predefined tables for short blocks, Appendix A.6), repeat-offset codes, optional XXH64 trailer;
// a raw block when that is not smaller.
//   match    CTA per chunk, the chunk staged in shared memory: each of the eight warps parses one eighth greedily with
//            its own hash table of 4-byte strings (first filled with every earlier position of the chunk, so that
//            candidates come from anywhere before): 32 positions probed per step, the first hit extended 128 bytes
//            per step by the whole warp; then the warps gather the literal runs from shared memory to their final
//            place in the frame
//   literals CTA per chunk: histogram, length-limited Huffman code (11 bits), tree description (FSE-coded weights or
//            the direct form), then the four streams in parallel, one warp each, 128 symbols per step at
//            prefix-summed bit offsets; raw or RLE literals when that is smaller
//   encode   warp per chunk: offsets -> repeat codes (forward, serial on one lane over coalesced loads), then the
//            backward FSE bitstream 32 sequences at a time: three lanes run the three state chains, every lane
//            assembles its sequence's bit fields at a prefix-summed bit offset in shared memory
//   finish   warp per chunk: raw-block fallback copy, XXH64 trailer, sizes
// kzg_enc_stage_encode (KZG_HD) is the serial statement of the encode stage: tests/hostsim runs it on the CPU behind
// a sequential matcher, and the GPU kernel must produce the same bytes for the same sequences.
#pragma once
#include "kzg_stages.cuh"

#define KZG_ENC_HASH_LOG 11          // per-warp table: 2048 entries of 16 bits
#define KZG_ENC_WARPS 8              // = parts per chunk (each warp parses one)
#define KZG_ENC_FLAG_CHECKSUM 1u
#define KZG_ENC_FLAG_LAZY 2u          // one step of lazy matching (better ratio, slower match stage)
#define KZG_ENC_ST_RAW 1u            // chunk state after encode: emit a raw block
#define KZG_ENC_ST_COMPRESSED 2u
#define KZG_ENC_ST_ERROR 3u

struct KzgEncTables {
    u16 st_ll[64], st_ml[64], st_of[32];              // FSE next-state tables of the predefined distributions
    u32 dnb_ll[36], dnb_ml[53], dnb_of[32];           // deltaNbBits
    i32 dfs_ll[36], dfs_ml[53], dfs_of[32];           // deltaFindState
    u32 ll_tab[36], ml_tab[53];                       // (extra_bits << 24) | baseline, as kzg_code_tables
    u8 ll_code[64], ml_code[128];                     // code of small literal lengths / (match length - 3)
};

struct KzgEncChunk { u32 n_seq[KZG_ENC_WARPS]; u32 lit_size, state, total, lit_section; };   // n_seq: sequences of each part; lit_section: bytes of the literals section in the block

struct KzgEncBatch {
    const u8* src; const u64* src_off; const u64* src_size;
    u8* dst; const u64* dst_off; const u64* dst_cap;
    u32* status; u64* out_size;
    u32 n, flags;
    u64* seq; u32 seq_cap;                            // records per chunk, an eighth of them per part: ll[0,18) | ml[18,36) | off[36,54)
    KzgEncChunk* chunks;
    u8* lit; u32 lit_cap;                             // literal bytes of chunk i, gathered by the match stage: lit + i * lit_cap
    const KzgEncTables* tab;
    u32* tickets;                                     // [0] match, [1] finish, [2] encode, [3] literals
};

KZG_HD void kzg_wr16(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); }
KZG_HD void kzg_wr24(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); }
KZG_HD void kzg_wr32(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24); }

// Frame header of a single-segment frame with content size n (Appendix A.1). Returns its length (6, 7 or 9).
KZG_HD u32 kzg_enc_header_size(u64 n) { return n < 256 ? 6u : (n < 65536 + 256 ? 7u : 9u); }
KZG_HD u32 kzg_enc_write_header(u8* p, u64 n, bool checksum) {
    kzg_wr32(p, KZG_MAGIC);
    const u32 flag = n < 256 ? 0u : (n < 65536 + 256 ? 1u : 2u);
    p[4] = (u8)((flag << 6) | (1u << 5) | (checksum ? 4u : 0u));
    if (flag == 0) { p[5] = (u8)n; return 6; }
    if (flag == 1) { kzg_wr16(p + 5, (u32)(n - 256)); return 7; }
    kzg_wr32(p + 5, (u32)n);
    return 9;
}
// Worst case of a frame for n input bytes: header, block header, raw content, checksum.
KZG_HD u64 kzg_enc_bound(u64 n) { return n + 32; }

// ---------------------------------------------------------------------------------------------
// FSE encoding tables of a normalised distribution (the mirror image of the decode-table build, Appendix A.5).
// scratch: 66 words + (1 << log) bytes.
KZG_HD void kzg_enc_build_ctable(const i16* norm, u32 nsym, u32 log, u16* state_table, u32* dnb, i32* dfs, u32* cumul, u8* symbol_of) {
    const u32 size = 1u << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    u32 high = size - 1;
    cumul[0] = 0;
    for (u32 s = 0; s < nsym; s++) {
        if (norm[s] == -1) { cumul[s + 1] = cumul[s] + 1; symbol_of[high--] = (u8)s; }
        else cumul[s + 1] = cumul[s] + (u32)norm[s];
    }
    u32 pos = 0;
    for (u32 s = 0; s < nsym; s++) {
        for (i32 k = 0; k < norm[s]; k++) {
            symbol_of[pos] = (u8)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    }
    for (u32 u = 0; u < size; u++) { const u32 s = symbol_of[u]; state_table[cumul[s]++] = (u16)(size + u); }
    u32 total = 0;
    for (u32 s = 0; s < nsym; s++) {
        const i32 c = norm[s];
        if (c == 0) { dnb[s] = ((log + 1) << 16) - size; dfs[s] = 0; }
        else if (c == -1 || c == 1) { dnb[s] = (log << 16) - size; dfs[s] = (i32)total - 1; total++; }
        else {
            const u32 max_bits = log - (u32)kzg_highbit((u32)c - 1);
            dnb[s] = (max_bits << 16) - ((u32)c << max_bits);
            dfs[s] = (i32)total - c;
            total += (u32)c;
        }
    }
}

inline void kzg_enc_build_tables(KzgEncTables& t) {
    KzgFseDesc d;
    kzg_default_fse_desc(d);
    memset(&t, 0, sizeof t);
    u32 cumul[66];
    u8 symbol_of[64];
    kzg_enc_build_ctable(d.norm[0], 36, 6, t.st_ll, t.dnb_ll, t.dfs_ll, cumul, symbol_of);
    kzg_enc_build_ctable(d.norm[1], 29, 5, t.st_of, t.dnb_of, t.dfs_of, cumul, symbol_of);
    kzg_enc_build_ctable(d.norm[2], 53, 6, t.st_ml, t.dnb_ml, t.dfs_ml, cumul, symbol_of);
    kzg_code_tables(t.ll_tab, t.ml_tab);
    for (u32 v = 0; v < 64; v++)
        for (u32 c = 0; c < 36; c++) { const u32 b = t.ll_tab[c] & 0xFFFFFFu, x = t.ll_tab[c] >> 24; if (v >= b && v < b + (1u << x)) t.ll_code[v] = (u8)c; }
    for (u32 v = 0; v < 128; v++)
        for (u32 c = 0; c < 53; c++) { const u32 b = t.ml_tab[c] & 0xFFFFFFu, x = t.ml_tab[c] >> 24; if (v + 3 >= b && v + 3 < b + (1u << x)) t.ml_code[v] = (u8)c; }
}

// ---------------------------------------------------------------------------------------------
// Forward bit writer (the decoder reads it backwards, Appendix A "Backward bitstreams").
struct KzgBitW { u64 acc; u32 nbits; u8* p; u8* end; bool ovf; };
KZG_HD void kzg_bw_add(KzgBitW& w, u32 v, u32 n) {          // n <= 31; keeps nbits < 64
    w.acc |= (u64)(v & ((1u << n) - 1u)) << w.nbits;
    w.nbits += n;
}
KZG_HD void kzg_bw_flush(KzgBitW& w) {
    while (w.nbits >= 8) {
        if (w.p < w.end) *w.p = (u8)w.acc; else w.ovf = true;
        w.p++;
        w.acc >>= 8; w.nbits -= 8;
    }
}
struct KzgFseC { u32 value; };
KZG_HD void kzg_fsec_init(KzgFseC& s, const u16* st, const u32* dnb, const i32* dfs, u32 sym) {
    const u32 nb = (dnb[sym] + (1u << 15)) >> 16;
    const u32 v = (nb << 16) - dnb[sym];
    s.value = st[(i32)(v >> nb) + dfs[sym]];
}
KZG_HD void kzg_fsec_encode(KzgBitW& w, KzgFseC& s, const u16* st, const u32* dnb, const i32* dfs, u32 sym) {
    const u32 nb = (s.value + dnb[sym]) >> 16;
    kzg_bw_add(w, s.value, nb);
    s.value = st[(i32)(s.value >> nb) + dfs[sym]];
}

KZG_HD u32 kzg_enc_ll_code(const KzgEncTables* t, u32 ll) { return ll < 64 ? (u32)t->ll_code[ll] : (u32)kzg_highbit(ll) + 19u; }
KZG_HD u32 kzg_enc_ml_code(const KzgEncTables* t, u32 mlbase) { return mlbase < 128 ? (u32)t->ml_code[mlbase] : (u32)kzg_highbit(mlbase) + 36u; }

// The three encoding tables one block is written with.  Index: 0 LL, 1 ML, 2 OF.
struct KzgEncTabSet {
    u16 st[1280];                   // state tables: LL at 0 (<= 512), ML at 512 (<= 512), OF at 1024 (<= 256)
    u32 dnb[128];                   // deltaNbBits: LL at 0, ML at 36, OF at 89
    i32 dfs[128];                   // deltaFindState, same layout
    u32 log[3];                     // accuracy log; 0 = RLE mode (the state carries no bits)
};
#define KZG_ENC_ST_OFF(t) ((t) == 0 ? 0u : ((t) == 1 ? 512u : 1024u))
#define KZG_ENC_SYM_OFF(t) ((t) == 0 ? 0u : ((t) == 1 ? 36u : 89u))
struct KzgEncTabScratch { u32 cumul[66]; i16 norm[64]; u8 symbol_of[512]; };

// the same on a table set: t = 0 LL, 1 ML, 2 OF
KZG_HD void kzg_fsec_init_set(KzgFseC& s, const KzgEncTabSet& ts, u32 t, u32 sym) {
    if (ts.log[t] == 0) { s.value = 1; return; }             // RLE mode
    kzg_fsec_init(s, ts.st + KZG_ENC_ST_OFF(t), ts.dnb + KZG_ENC_SYM_OFF(t), ts.dfs + KZG_ENC_SYM_OFF(t), sym);
}
KZG_HD void kzg_fsec_encode_set(KzgBitW& w, KzgFseC& s, const KzgEncTabSet& ts, u32 t, u32 sym) {
    kzg_fsec_encode(w, s, ts.st + KZG_ENC_ST_OFF(t), ts.dnb + KZG_ENC_SYM_OFF(t), ts.dfs + KZG_ENC_SYM_OFF(t), sym);
}

// Counts -> normalised counts that sum to 1 << log; every present symbol keeps at least 1 (Appendix A.5).
KZG_HD void kzg_enc_normalize(const u32* cnt, u32 nsym, u32 total, u32 log, i16* norm) {
    const u32 target = 1u << log;
    u32 sum = 0, imax = 0, pmax = 0;
    for (u32 s = 0; s < nsym; s++) {
        u32 pr = 0;
        if (cnt[s]) { pr = (u32)((((u64)cnt[s] << log) + (total >> 1)) / total); if (pr == 0) pr = 1; }
        norm[s] = (i16)pr;
        sum += pr;
        if (pr > pmax) { pmax = pr; imax = s; }
    }
    if (sum < target) norm[imax] = (i16)(norm[imax] + (target - sum));
    while (sum > target) {                                  // take the excess from the largest entries
        u32 im = 0, pm = 0;
        for (u32 s = 0; s < nsym; s++) if ((u32)norm[s] > pm) { pm = (u32)norm[s]; im = s; }
        const u32 d = pm - 1 < sum - target ? pm - 1 : sum - target;
        norm[im] = (i16)(pm - d);
        sum -= d;
    }
}

// FSE table description (Appendix A.5), the inverse of kzg_read_ncount.  Returns bytes written, 0 if they do not fit.
KZG_HD u32 kzg_enc_write_ncount(u8* dst, u32 cap, const i16* norm, u32 nsym, u32 log) {
    KzgBitW w;
    w.acc = 0; w.nbits = 0; w.p = dst; w.end = dst + cap; w.ovf = false;
    kzg_bw_add(w, log - 5, 4);
    i32 remaining = 1 << log;
    u32 s = 0;
    while (remaining > 0 && s < nsym) {
        const u32 bits = (u32)kzg_highbit((u32)remaining + 1) + 1;
        const u32 lower = (1u << (bits - 1)) - 1, thr = (1u << bits) - 1 - ((u32)remaining + 1);
        const u32 val = (u32)(norm[s] + 1);
        if (val < thr) kzg_bw_add(w, val, bits - 1);
        else if (val <= lower) kzg_bw_add(w, val, bits);
        else kzg_bw_add(w, val + thr, bits);
        remaining -= norm[s];
        const bool zero = norm[s] == 0;
        s++;
        if (zero) {                                          // how many more zero-probability symbols follow
            u32 n0 = 0;
            while (s + n0 < nsym && norm[s + n0] == 0) n0++;
            s += n0;
            while (n0 >= 3) { kzg_bw_add(w, 3, 2); n0 -= 3; kzg_bw_flush(w); }
            kzg_bw_add(w, n0, 2);
        }
        kzg_bw_flush(w);
    }
    if (w.nbits) { if (w.p < w.end) *w.p = (u8)w.acc; else w.ovf = true; w.p++; }
    return (w.ovf || remaining != 0) ? 0u : (u32)(w.p - dst);
}

// Chooses the mode of each sequence table from the code histograms (hist: LL at 0, ML at 36, OF at 89), writes the
// modes byte (at p[-1]) and the table descriptions at p, fills `ts`.  Returns the new write position (NULL: no room).
KZG_HD u8* kzg_enc_prepare_tables(const KzgEncTables* T, u32 nseq, const u32* hist, KzgEncTabSet& ts, KzgEncTabScratch& sc, u8* p, u8* limit) {
    u8* const modes_byte = p - 1;
    const u32 nsym_of[3] = { 36, 53, 32 }, max_log[3] = { 9, 9, 8 }, pre_log[3] = { 6, 6, 5 };
    u32 mode[3];
    // descriptions are stored in the order LL, OF, ML (Appendix A.6)
    const u32 order[3] = { 0, 2, 1 };
    for (u32 k = 0; k < 3; k++) {
        const u32 t = order[k];
        const u32* h = hist + KZG_ENC_SYM_OFF(t);
        u16* st = ts.st + KZG_ENC_ST_OFF(t);
        u32* dnb = ts.dnb + KZG_ENC_SYM_OFF(t);
        i32* dfs = ts.dfs + KZG_ENC_SYM_OFF(t);
        u32 present = 0, last = 0;
        for (u32 s = 0; s < nsym_of[t]; s++) if (h[s]) { present++; last = s; }
        if (present == 1 && nseq >= 8) {                     // RLE mode: one byte, no state bits
            mode[t] = 1;
            if (p >= limit) return nullptr;
            *p++ = (u8)last;
            ts.log[t] = 0;
            st[0] = 1; dnb[last] = 0; dfs[last] = -1;
        } else if (nseq >= 128 && present > 1) {             // FSE mode: a table made for this block
            mode[t] = 2;
            u32 log = (u32)kzg_highbit(nseq - 1) - 2;
            const u32 min_log = (u32)kzg_highbit(last) + 2;
            if (log < min_log) log = min_log;
            if (log < 5) log = 5;
            if (log > max_log[t]) log = max_log[t];
            kzg_enc_normalize(h, last + 1, nseq, log, sc.norm);
            const u32 nb = kzg_enc_write_ncount(p, (u32)(limit - p), sc.norm, last + 1, log);
            if (nb == 0) return nullptr;
            p += nb;
            ts.log[t] = log;
            kzg_enc_build_ctable(sc.norm, last + 1, log, st, dnb, dfs, sc.cumul, sc.symbol_of);
        } else {                                             // predefined
            mode[t] = 0;
            ts.log[t] = pre_log[t];
            const u16* pst = t == 0 ? T->st_ll : (t == 1 ? T->st_ml : T->st_of);
            const u32* pdnb = t == 0 ? T->dnb_ll : (t == 1 ? T->dnb_ml : T->dnb_of);
            const i32* pdfs = t == 0 ? T->dfs_ll : (t == 1 ? T->dfs_ml : T->dfs_of);
            for (u32 u = 0; u < (1u << pre_log[t]); u++) st[u] = pst[u];
            for (u32 s = 0; s < nsym_of[t]; s++) { dnb[s] = pdnb[s]; dfs[s] = pdfs[s]; }
        }
    }
    *modes_byte = (u8)((mode[0] << 6) | (mode[2] << 4) | (mode[1] << 2));     // LL, OF, ML
    return p;
}

// FSE-coded Huffman weights (Appendix A.4, tree description with a header byte < 128): n weights (the last symbol's
// is implicit and not part of them), alphabet 0..12, accuracy log 5 or 6, two interleaved states.  Writes the table
// description and the bitstream to out (not the header byte); returns their size, or 0 when that does not work out
// (a single distinct weight, no room): the caller then falls back to the direct 4-bit form.
struct KzgEncWeightScratch { u16 st[64]; u32 dnb[16]; i32 dfs[16]; u32 cnt[16]; KzgEncTabScratch tab; };
KZG_HD u32 kzg_enc_compress_weights(const u8* w, u32 n, u8* out, u32 cap, KzgEncWeightScratch& sc) {
    if (n < 2) return 0;
    u32 maxw = 0, distinct = 0;
    for (u32 k = 0; k < 16; k++) sc.cnt[k] = 0;
    for (u32 k = 0; k < n; k++) { sc.cnt[w[k]]++; if (w[k] > maxw) maxw = w[k]; }
    for (u32 k = 0; k <= maxw; k++) distinct += sc.cnt[k] != 0;
    if (distinct < 2) return 0;
    u32 log = 6;
    const u32 src_bits = (u32)kzg_highbit(n - 1) < 2 ? 0 : (u32)kzg_highbit(n - 1) - 2;
    if (src_bits < log) log = src_bits;
    u32 min_bits = (u32)kzg_highbit(n) + 1;
    if (maxw && (u32)kzg_highbit(maxw) + 2 < min_bits) min_bits = (u32)kzg_highbit(maxw) + 2;
    if (log < min_bits) log = min_bits;
    if (log < 5) log = 5;
    if (log > 6) return 0;
    kzg_enc_normalize(sc.cnt, maxw + 1, n, log, sc.tab.norm);
    const u32 hsz = kzg_enc_write_ncount(out, cap, sc.tab.norm, maxw + 1, log);
    if (hsz == 0) return 0;
    kzg_enc_build_ctable(sc.tab.norm, maxw + 1, log, sc.st, sc.dnb, sc.dfs, sc.tab.cumul, sc.tab.symbol_of);
    KzgBitW bw;
    bw.acc = 0; bw.nbits = 0; bw.p = out + hsz; bw.end = out + cap; bw.ovf = false;
    KzgFseC s1, s2;                                          // state 1 carries the even positions, state 2 the odd ones
    u32 k = n;
    if (n & 1) {
        kzg_fsec_init(s1, sc.st, sc.dnb, sc.dfs, w[--k]);
        kzg_fsec_init(s2, sc.st, sc.dnb, sc.dfs, w[--k]);
        kzg_fsec_encode(bw, s1, sc.st, sc.dnb, sc.dfs, w[--k]);
        kzg_bw_flush(bw);
    } else {
        kzg_fsec_init(s2, sc.st, sc.dnb, sc.dfs, w[--k]);
        kzg_fsec_init(s1, sc.st, sc.dnb, sc.dfs, w[--k]);
    }
    while (k > 0) {
        kzg_fsec_encode(bw, s2, sc.st, sc.dnb, sc.dfs, w[--k]);
        kzg_fsec_encode(bw, s1, sc.st, sc.dnb, sc.dfs, w[--k]);
        kzg_bw_flush(bw);
    }
    kzg_bw_add(bw, s2.value, log);
    kzg_bw_add(bw, s1.value, log);
    kzg_bw_add(bw, 1, 1);                                    // end mark
    kzg_bw_flush(bw);
    if (bw.nbits) { if (bw.p < bw.end) *bw.p = (u8)bw.acc; else bw.ovf = true; bw.p++; }
    return bw.ovf ? 0u : (u32)(bw.p - out);
}

// Repeat-offset code of a distance (Appendix A.7); updates the history like the decoder will.
KZG_HD u32 kzg_enc_offset_value(u32 ll, u32 off, u32& r0, u32& r1, u32& r2) {
    u32 ob;
    if (ll) {
        if (off == r0) return 1;
        if (off == r1) { r1 = r0; r0 = off; return 2; }
        ob = off == r2 ? 3u : off + 3u;
    } else {
        if (off == r1) { r1 = r0; r0 = off; return 1; }
        ob = off == r2 ? 2u : (off == r0 - 1 ? 3u : off + 3u);
    }
    r2 = r1; r1 = r0; r0 = off;
    return ob;
}

// Header of a raw (type 0) or RLE (type 1) literals section with a 20-bit size: 3 bytes (Appendix A.3).
KZG_HD void kzg_enc_write_lit_header(u8* blk, u32 type, u32 lit_size) {
    blk[0] = (u8)(type | (3u << 2) | ((lit_size & 15u) << 4));
    blk[1] = (u8)(lit_size >> 4);
    blk[2] = (u8)(lit_size >> 12);
}

// Headers in front of the sequence bitstream. Returns where the bitstream starts, or NULL (raw block / error:
// ck.state says which).
KZG_HD u8* kzg_enc_block_prologue(const KzgEncBatch& B, u32 i, u32 nseq, u8*& blk, u8*& limit) {
    const u64 n = B.src_size[i];
    u8* const dst = B.dst + B.dst_off[i];
    KzgEncChunk& ck = B.chunks[i];
    ck.total = 0;
    if (n > KZG_BLOCK_MAX) { ck.state = KZG_ENC_ST_ERROR; B.status[i] = KZG_E_SRCSIZE_WRONG; B.out_size[i] = 0; return nullptr; }
    if (B.dst_cap[i] < kzg_enc_bound(n)) { ck.state = KZG_ENC_ST_ERROR; B.status[i] = KZG_E_DSTSIZE_TOO_SMALL; B.out_size[i] = 0; return nullptr; }
    const u32 hdr = kzg_enc_write_header(dst, n, (B.flags & KZG_ENC_FLAG_CHECKSUM) != 0);
    ck.state = KZG_ENC_ST_RAW;
    if (nseq == 0 || ck.lit_section == 0) return nullptr;
    blk = dst + hdr + 3;
    u8* p = blk + ck.lit_section;                                // the literals section is in place already
    limit = blk + n;                                             // a compressed block must be smaller than the raw one
    if (p + 4 >= limit) return nullptr;
    if (nseq < 128) *p++ = (u8)nseq;
    else if (nseq < 0x7F00) { *p++ = (u8)((nseq >> 8) + 0x80); *p++ = (u8)nseq; }
    else { *p++ = 0xFF; kzg_wr16(p, nseq - 0x7F00); p += 2; }
    *p++ = 0;                                                    // modes byte (kzg_enc_prepare_tables fills it in)
    return p;
}
KZG_HD void kzg_enc_block_epilogue(const KzgEncBatch& B, u32 i, u8* blk, u8* end) {
    u8* const dst = B.dst + B.dst_off[i];
    const u32 hdr = (u32)(blk - dst) - 3, bsize = (u32)(end - blk);
    kzg_wr24(dst + hdr, 1u | (2u << 1) | (bsize << 3));
    B.chunks[i].state = KZG_ENC_ST_COMPRESSED;
    B.chunks[i].total = hdr + 3 + bsize;
}

// Record s (in order) of a chunk whose sequences sit in KZG_ENC_WARPS regions of seq_cap / KZG_ENC_WARPS records.
KZG_HD u64* kzg_enc_seq_at(u64* seq, u32 qcap, const u32* nq, u32 s) {
    u32 q = 0;
    while (s >= nq[q]) { s -= nq[q]; q++; }
    return seq + (size_t)q * qcap + s;
}

// encode stage, serial statement: one thread per chunk.  The match stage has left: literals at dst + header + 6,
// records, counts.
KZG_HD void kzg_enc_stage_encode(const KzgEncBatch& B, u32 i) {
    KzgEncChunk& ck = B.chunks[i];
    u32 nq[KZG_ENC_WARPS], nseq = 0;
    for (u32 q = 0; q < KZG_ENC_WARPS; q++) { nq[q] = ck.n_seq[q]; nseq += nq[q]; }
    const u32 qcap = B.seq_cap / KZG_ENC_WARPS;
    u8 *blk = nullptr, *limit = nullptr;
    u8* p = kzg_enc_block_prologue(B, i, nseq, blk, limit);
    if (!p) return;
    const KzgEncTables* T = B.tab;
    u64* seq = B.seq + (size_t)i * B.seq_cap;
    u32 hist[128];
    for (u32 k = 0; k < 128; k++) hist[k] = 0;
    {   // forward: distances -> offset values with repeat codes, in place; code histograms
        u32 r0 = 1, r1 = 4, r2 = 8;
        for (u32 s = 0; s < nseq; s++) {
            u64* r = kzg_enc_seq_at(seq, qcap, nq, s);
            const u64 rec = *r;
            const u32 ll = (u32)rec & 0x3FFFFu, mlb = ((u32)(rec >> 18) & 0x3FFFFu) - 3u;
            const u32 ob = kzg_enc_offset_value(ll, (u32)(rec >> 36) & 0x3FFFFu, r0, r1, r2);
            *r = (rec & 0xFFFFFFFFFull) | ((u64)ob << 36);
            hist[kzg_enc_ll_code(T, ll)]++; hist[36 + kzg_enc_ml_code(T, mlb)]++; hist[89 + (u32)kzg_highbit(ob)]++;
        }
    }
    KzgEncTabSet ts;
    KzgEncTabScratch sc;
    p = kzg_enc_prepare_tables(T, nseq, hist, ts, sc, p, limit);
    if (!p) return;
    // backward: the last sequence initialises the states, the first is written last
    KzgBitW w;
    w.acc = 0; w.nbits = 0; w.p = p; w.end = limit; w.ovf = false;
    KzgFseC sLL, sML, sOF;
    {
        const u64 rec = *kzg_enc_seq_at(seq, qcap, nq, nseq - 1);
        const u32 ll = (u32)rec & 0x3FFFFu, mlb = ((u32)(rec >> 18) & 0x3FFFFu) - 3u, ob = (u32)(rec >> 36) & 0x3FFFFu;
        const u32 cl = kzg_enc_ll_code(T, ll), cm = kzg_enc_ml_code(T, mlb), co = (u32)kzg_highbit(ob);
        kzg_fsec_init_set(sML, ts, 1, cm);
        kzg_fsec_init_set(sOF, ts, 2, co);
        kzg_fsec_init_set(sLL, ts, 0, cl);
        kzg_bw_add(w, ll, T->ll_tab[cl] >> 24);
        kzg_bw_add(w, mlb, T->ml_tab[cm] >> 24);
        kzg_bw_flush(w);
        kzg_bw_add(w, ob, co);
        kzg_bw_flush(w);
    }
    for (u32 s = nseq - 1; s-- > 0;) {
        const u64 rec = *kzg_enc_seq_at(seq, qcap, nq, s);
        const u32 ll = (u32)rec & 0x3FFFFu, mlb = ((u32)(rec >> 18) & 0x3FFFFu) - 3u, ob = (u32)(rec >> 36) & 0x3FFFFu;
        const u32 cl = kzg_enc_ll_code(T, ll), cm = kzg_enc_ml_code(T, mlb), co = (u32)kzg_highbit(ob);
        kzg_fsec_encode_set(w, sOF, ts, 2, co);
        kzg_fsec_encode_set(w, sML, ts, 1, cm);
        kzg_fsec_encode_set(w, sLL, ts, 0, cl);
        kzg_bw_add(w, ll, T->ll_tab[cl] >> 24);
        kzg_bw_flush(w);
        kzg_bw_add(w, mlb, T->ml_tab[cm] >> 24);
        kzg_bw_add(w, ob, co);
        kzg_bw_flush(w);
        if (w.ovf) return;                                       // not smaller than raw
    }
    kzg_bw_add(w, sML.value, ts.log[1]);
    kzg_bw_add(w, sOF.value, ts.log[2]);
    kzg_bw_add(w, sLL.value, ts.log[0]);
    kzg_bw_add(w, 1, 1);                                         // end mark
    kzg_bw_flush(w);
    if (w.nbits) { if (w.p < w.end) *w.p = (u8)w.acc; else w.ovf = true; w.p++; }
    if (w.ovf || w.p >= limit) return;
    kzg_enc_block_epilogue(B, i, blk, w.p);
}

#if defined(__CUDACC__)
namespace kzg_encd {

__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ u32 lds8(u32 a) { u32 v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
// unaligned 32-bit read from shared memory (two aligned words)
__device__ __forceinline__ u32 lds32u(u32 a) {
    const u32 b = a & ~3u;
    return __funnelshift_r(lds32(b), lds32(b + 4), (a & 3u) << 3);
}
// the 4 bytes at shared address a and the table slot of the 5 bytes there (longer context: candidates that are more
// likely to be long matches)
__device__ __forceinline__ u32 lds32u_hash5(u32 a, u32& h) {
    const u32 b = a & ~3u, sh = (a & 3u) << 3;
    const u32 w0 = lds32(b), w1 = lds32(b + 4), w2 = lds32(b + 8);
    const u32 v = __funnelshift_r(w0, w1, sh);
    h = ((v * 2654435761u) ^ ((__funnelshift_r(w1, w2, sh) & 0xFFFFu) * 0x9E3779B1u)) * 0x85EBCA6Bu >> (32 - KZG_ENC_HASH_LOG);
    return v;
}
// unaligned 32-bit read from global memory (two aligned words; reads up to 7 bytes past p)
__device__ __forceinline__ u32 rd32u(const u8* p) {
    const u32* q = reinterpret_cast<const u32*>((uintptr_t)p & ~(uintptr_t)3);
    const u32 sh = ((u32)(uintptr_t)p & 3u) << 3;
    const u32 a = q[0];
    if (sh == 0) return a;
    return __funnelshift_r(a, q[1], sh);
}
// warp copy of n bytes between global buffers, any alignment; reads stay inside [src & ~3, src + n + 3]
__device__ __forceinline__ void warp_copy(u8* dst, const u8* src, u32 n, u32 lane) {
    if (n < 64) { for (u32 k = lane; k < n; k += 32) dst[k] = src[k]; return; }
    const u32 head = (u32)((4 - ((uintptr_t)dst & 3)) & 3);
    if (lane < head) dst[lane] = src[lane];
    const u32 nw = (n - head) >> 2;
    u32* d4 = reinterpret_cast<u32*>(dst + head);
    const u8* s = src + head;
    for (u32 k = lane; k < nw; k += 32) d4[k] = rd32u(s + 4 * k);
    for (u32 k = head + (nw << 2) + lane; k < n; k += 32) dst[k] = src[k];
}
// the same for a range that ends at the end of a chunk: the last bytes go one at a time, because warp_copy's
// word reads may touch up to 3 bytes past its range
__device__ __forceinline__ void warp_copy_to_end(u8* dst, const u8* src, u32 n, u32 lane) {
    if (n < 72) { for (u32 k = lane; k < n; k += 32) dst[k] = src[k]; return; }
    warp_copy(dst, src, n - 8, lane);
    if (lane < 8) dst[n - 8 + lane] = src[n - 8 + lane];
}
// warp copy of n bytes from shared memory (address sa) to global memory
__device__ __forceinline__ void warp_copy_from_smem(u8* dst, u32 sa, u32 n, u32 lane) {
    if (n < 64) { for (u32 k = lane; k < n; k += 32) dst[k] = (u8)lds8(sa + k); return; }
    const u32 head = (u32)((4 - ((uintptr_t)dst & 3)) & 3);
    if (lane < head) dst[lane] = (u8)lds8(sa + lane);
    const u32 nw = (n - head) >> 2;
    u32* d4 = reinterpret_cast<u32*>(dst + head);
    for (u32 k = lane; k < nw; k += 32) d4[k] = lds32u(sa + head + 4 * k);
    for (u32 k = head + (nw << 2) + lane; k < n; k += 32) dst[k] = (u8)lds8(sa + k);
}
__device__ __forceinline__ u32 warp_incl_sum(u32 v, u32 lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= (u32)o) v += t; }
    return v;
}

}  // namespace kzg_encd

// Shared memory of the match kernel: [chunk area: chunk_cap + 32 bytes][one hash table per warp][misc: 128 bytes]
#define KZG_ENC_TABLE_BYTES (2u << KZG_ENC_HASH_LOG)
KZG_HD u32 kzg_enc_match_smem(u32 chunk_cap) { return chunk_cap + 32u + KZG_ENC_WARPS * KZG_ENC_TABLE_BYTES + 128u; }

// match stage: one CTA per chunk.
__global__ void __launch_bounds__(KZG_ENC_WARPS * 32) kzg_enc_match_kernel(KzgEncBatch B, u32 chunk_cap) {
    using namespace kzg_encd;
    extern __shared__ uint4 enc_smem4[];
    u8* const sm = reinterpret_cast<u8*>(enc_smem4);
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u16* const table = reinterpret_cast<u16*>(sm + chunk_cap + 32) + (warp << KZG_ENC_HASH_LOG);
    // misc: [0,8) sequences, [8,16) literal bytes, [16,24) tail literals of the parts, [24] ticket
    u32* const misc = reinterpret_cast<u32*>(sm + chunk_cap + 32 + KZG_ENC_WARPS * KZG_ENC_TABLE_BYTES);
    const u32 s_chunk = smem_addr(sm);
    const bool lazy = (B.flags & KZG_ENC_FLAG_LAZY) != 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) misc[24] = atomicAdd(&B.tickets[0], 1u);
        __syncthreads();
        const u32 i = misc[24];
        if (i >= B.n) break;
        const u32 n = (u32)B.src_size[i];
        if (B.src_size[i] > chunk_cap || B.src_size[i] > KZG_BLOCK_MAX || B.dst_cap[i] < kzg_enc_bound(n) || n < 64) {   // errors / tiny chunks: raw block
            if (threadIdx.x == 0) {
                KzgEncChunk ck;
                for (u32 q = 0; q < KZG_ENC_WARPS; q++) ck.n_seq[q] = 0;
                ck.lit_size = 0; ck.state = 0; ck.total = 0; ck.lit_section = 0;
                B.chunks[i] = ck;
            }
            continue;
        }
        // ---- stage the chunk, keeping its 16-byte phase
        const u8* g = B.src + B.src_off[i];
        const u32 shift = (u32)((uintptr_t)g & 15);
        {
            const uint4* g4 = reinterpret_cast<const uint4*>(g - shift);
            const u32 nvec = (shift + n + 15) >> 4;
            for (u32 v = threadIdx.x; v < nvec; v += blockDim.x) enc_smem4[v] = __ldg(g4 + v);
        }
        __syncthreads();
        const u32 base = s_chunk + shift;                       // shared address of chunk byte 0
        // what hashing may read beyond the chunk is fixed, so that the output depends on the chunk alone
        if (threadIdx.x < 16) sm[shift + n + threadIdx.x] = 0;
        __syncthreads();
        // ---- this warp's part
        const u32 qlen = (((n + KZG_ENC_WARPS - 1) / KZG_ENC_WARPS) + 31u) & ~31u;
        const u32 q0 = min(warp * qlen, n), q1 = min(q0 + qlen, n);
        const u32 qcap = B.seq_cap / KZG_ENC_WARPS;
        u64* const seqq = B.seq + (size_t)i * B.seq_cap + (size_t)warp * qcap;
        // its table starts as "every earlier position of the chunk, inserted in order" (positions keep their low 16 bits)
        for (u32 k = lane; k < (1u << KZG_ENC_HASH_LOG) / 8; k += 32) reinterpret_cast<uint4*>(table)[k] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        for (u32 p0 = 0; p0 < q0; p0 += 32) {                   // q0 is a multiple of 32
            const u32 p = p0 + lane;
            u32 h;
            lds32u_hash5(base + p, h);
            table[h] = (u16)p;
        }
        __syncwarp();
        u32 pos = q0, anchor = q0, nseq = 0, nlit = 0, rep = 0xFFFFFFFFu;
        const u32 limit = q1 >= 3 ? q1 - 3 : 0;                  // probed positions have their 4 bytes inside the part
        while (pos < limit) {
            const u32 p = pos + lane;
            const bool ok = p < limit;
            u32 h;
            const u32 v = lds32u_hash5(base + p, h);
            // the table keeps the low 16 bits of positions: the candidate is the nearest earlier position with them
            i32 cand = (i32)((p & ~0xFFFFu) | (u32)table[h]);
            if (cand >= (i32)p) cand -= 65536;
            // the distance of the previous match first: it costs a repeat code instead of a full offset
            const bool rep_hit = ok && p >= rep && lds32u(base + p - rep) == v;
            if (rep_hit) cand = (i32)(p - rep);
            const bool hit = rep_hit || (ok && cand >= 0 && lds32u(base + (u32)cand) == v);
            const u32 m = __ballot_sync(0xFFFFFFFFu, hit);
            if (m == 0) {                                        // nothing here: all 32 positions enter the table
                if (ok) table[h] = (u16)p;
                __syncwarp();
                pos += 32;
                continue;
            }
            // Every hit of this step that starts at or behind the end of the previous match becomes a sequence (short
            // matches are the rule: probing again for each of them would cost more than it finds).  Positions outside
            // the matches enter the table; the ones behind the last match are probed again by the next step and must
            // still find their older candidates then.
            u32 avail = m, ins = 0, from = 0;                    // hits still usable; lanes to insert; first lane not yet covered
            bool full = false;
            // length of the match between positions a and b (b < a) beyond their first `have` bytes, inside the part
            auto extend = [&](u32 a, u32 b, u32 have) {
                u32 len = have;
                for (;;) {                                       // 128 bytes per step
                    const u32 k = len + 4 * lane;
                    const i32 rem = (i32)q1 - (i32)(a + k);      // bytes of this lane's word that belong to the part
                    u32 diff = 1u;
                    if (rem > 0) {
                        diff = lds32u(base + a + k) ^ lds32u(base + b + k);
                        if (rem < 4) diff = (diff & ((1u << (8 * rem)) - 1u)) | (1u << (8 * rem));
                    }
                    const u32 mm = __ballot_sync(0xFFFFFFFFu, diff != 0);
                    if (mm) {
                        const u32 fl = __ffs(mm) - 1;
                        const u32 d = __shfl_sync(0xFFFFFFFFu, diff, fl);
                        return len + 4 * fl + (((u32)__ffs(d) - 1) >> 3);
                    }
                    len += 128;
                }
            };
            while (avail) {
                u32 l = __ffs(avail) - 1;
                u32 ms = pos + l, mc = (u32)__shfl_sync(0xFFFFFFFFu, cand, l);
                u32 len = extend(ms, mc, 4);
                // one step of lazy matching: a hit at the next position that is longer by more than the literal it costs
                if (lazy && l < 31 && ((avail >> (l + 1)) & 1u)) {
                    const u32 mc2 = (u32)__shfl_sync(0xFFFFFFFFu, cand, l + 1);
                    const u32 len2 = extend(ms + 1, mc2, 4);
                    if (len2 > len + 1) { l++; ms++; mc = mc2; len = len2; }
                }
                ins |= (l == 31 ? 0xFFFFFFFFu : ((2u << l) - 1u)) & ~((1u << from) - 1u);      // lanes from .. l
                {   // backward extension into the pending literals (at most 32 bytes)
                    const bool same = ms > anchor + lane && mc > lane && lds8(base + ms - 1 - lane) == lds8(base + mc - 1 - lane);
                    const u32 stop = __ballot_sync(0xFFFFFFFFu, !same);
                    const u32 back = stop ? __ffs(stop) - 1 : 32u;
                    ms -= back; mc -= back; len += back;
                }
                const u32 ll = ms - anchor;
                if (lane == 0) seqq[nseq] = (u64)ll | ((u64)len << 18) | ((u64)(ms - mc) << 36);
                nseq++;
                nlit += ll;
                anchor = ms + len;
                rep = ms - mc;
                if (nseq == qcap) { full = true; break; }
                const u32 end_rel = anchor - pos;                // first position behind the match, relative to this step
                if (end_rel >= 32) break;
                from = end_rel;
                // the distance of this match is the repeat distance for what follows: probe it for the positions left
                const bool again = ok && lane >= end_rel && p >= rep && lds32u(base + p - rep) == v;
                if (again) cand = (i32)(p - rep);
                avail = (avail | __ballot_sync(0xFFFFFFFFu, again)) & ~((1u << end_rel) - 1u);
            }
            if (ok && ((ins >> lane) & 1u)) table[h] = (u16)p;
            __syncwarp();
            pos = anchor;
            if (full) break;
        }
        const u32 tail = q1 - anchor;
        if (lane == 0) { misc[warp] = nseq; misc[8 + warp] = nlit + tail; misc[16 + warp] = tail; }
        __syncthreads();
        // ---- literal runs of this part, from shared memory to their final place
        u32 lbase = 0, carry = 0;                                // literal bytes before the part; literals pending in front of it
        for (u32 q = 0; q < warp; q++) {
            lbase += misc[8 + q];
            carry = misc[q] ? misc[16 + q] : carry + misc[8 + q];
        }
        u8* const lit = B.lit + (size_t)i * B.lit_cap + lbase;
        u32 spos = q0, dpos = 0;
        for (u32 b = 0; b < nseq; b += 32) {
            const bool have = b + lane < nseq;
            const u64 rec = have ? seqq[b + lane] : 0ull;
            const u32 ll = (u32)rec & 0x3FFFFu, ml = (u32)(rec >> 18) & 0x3FFFFu;
            const u32 sa = warp_incl_sum(ll + ml, lane), sl = warp_incl_sum(ll, lane);
            const u32 my_src = spos + sa - ll - ml, my_dst = dpos + sl - ll;
            // short runs: every lane copies its own; long ones: the warp together
            const u32 longm = __ballot_sync(0xFFFFFFFFu, ll >= 64);
            if (ll < 64) for (u32 t = 0; t < ll; t++) lit[my_dst + t] = (u8)lds8(base + my_src + t);
            for (u32 mm = longm; mm; mm &= mm - 1) {
                const u32 j = __ffs(mm) - 1;
                warp_copy_from_smem(lit + __shfl_sync(0xFFFFFFFFu, my_dst, j), base + __shfl_sync(0xFFFFFFFFu, my_src, j),
                                    __shfl_sync(0xFFFFFFFFu, ll, j), lane);
            }
            spos += __shfl_sync(0xFFFFFFFFu, sa, 31);
            dpos += __shfl_sync(0xFFFFFFFFu, sl, 31);
        }
        warp_copy_from_smem(lit + dpos, base + q1 - tail, tail, lane);
        // literals pending in front of the part belong to its first sequence
        if (lane == 0 && nseq && carry) seqq[0] += carry;
        if (threadIdx.x == 0) {
            KzgEncChunk ck;
            ck.lit_size = 0;
            for (u32 q = 0; q < KZG_ENC_WARPS; q++) { ck.n_seq[q] = misc[q]; ck.lit_size += misc[8 + q]; }
            ck.state = 0; ck.total = 0; ck.lit_section = 0;
            B.chunks[i] = ck;
        }
    }
}

// literals stage: one CTA per chunk.  Reads the gathered literals from the workspace and writes the literals section
// (Appendix A.3-A.4) to its place in the block: Huffman-coded in four streams, or raw / RLE.
#define KZG_ENC_HUF_THREADS 128
__global__ void __launch_bounds__(KZG_ENC_HUF_THREADS) kzg_enc_literals_kernel(KzgEncBatch B) {
    using namespace kzg_encd;
    __shared__ u32 hist[256];
    __shared__ u32 wnode[512];                     // weights of the tree nodes (leaves in sorted order, then internal)
    __shared__ u16 parent[512];
    __shared__ u16 order[256];                     // present symbols by ascending count
    __shared__ u8 len[256];                        // code length per symbol, 0 = absent
    __shared__ u16 code[256];
    __shared__ u32 stage[4][48];                   // per stream: 128 symbols x 11 bits + carry
    __shared__ KzgEncWeightScratch wsc;            // FSE-coded weight table (thread 0)
    __shared__ u8 wts[256], wbuf[128];
    __shared__ u32 misc[16];                       // [0] ticket, [1] max symbol, [2] mode, [3] maxbits, [4..8) stream bits, [8..12) stream offsets, [12] header bytes
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        __syncthreads();
        if (tid == 0) { misc[0] = atomicAdd(&B.tickets[3], 1u); misc[1] = 0; }
        for (u32 k = tid; k < 256; k += blockDim.x) { hist[k] = 0; len[k] = 0; }
        __syncthreads();
        const u32 i = misc[0];
        if (i >= B.n) break;
        const KzgEncChunk ck = B.chunks[i];
        u32 nseq = 0;
#pragma unroll
        for (int q = 0; q < KZG_ENC_WARPS; q++) nseq += ck.n_seq[q];
        if (nseq == 0) continue;                                 // raw block anyway
        const u32 n = (u32)B.src_size[i], nl = ck.lit_size;
        const u8* lit = B.lit + (size_t)i * B.lit_cap;
        u8* const blk = B.dst + B.dst_off[i] + kzg_enc_header_size(n) + 3;
        // ---- histogram
        {
            const u32* lit4 = reinterpret_cast<const u32*>(lit);
            const u32 nw = nl >> 2;
            for (u32 k = tid; k < nw; k += blockDim.x) {
                const u32 w = lit4[k];
                atomicAdd(&hist[w & 255u], 1u); atomicAdd(&hist[(w >> 8) & 255u], 1u);
                atomicAdd(&hist[(w >> 16) & 255u], 1u); atomicAdd(&hist[w >> 24], 1u);
            }
            for (u32 k = (nw << 2) + tid; k < nl; k += blockDim.x) atomicAdd(&hist[lit[k]], 1u);
        }
        __syncthreads();
        u32 present = 0;
        for (u32 k = tid; k < 256; k += blockDim.x) if (hist[k]) { present++; atomicMax(&misc[1], k); }
        const u32 m = (u32)__syncthreads_count(present == 1) + 2u * (u32)__syncthreads_count(present == 2);
        const u32 maxsym = misc[1];
        // mode: 0 raw, 1 RLE, 2 Huffman
        u32 mode = 2;
        if (nl < 64) mode = 0;
        if (m == 1 && nl > 0) mode = 1;
        if (mode == 2) {
            // ---- present symbols by ascending count (rank sort), then the Huffman tree on one thread
            for (u32 t = tid; t < 256; t += blockDim.x) {
                const u32 c = hist[t];
                if (c) {
                    u32 r = 0;
                    for (u32 j = 0; j < 256; j++) { const u32 cj = hist[j]; r += (cj != 0) && (cj < c || (cj == c && j < t)); }
                    order[r] = (u16)t;
                }
            }
            __syncthreads();
            if (tid == 0) {
                for (u32 k = 0; k < m; k++) wnode[k] = hist[order[k]];
                u32 maxd;
                for (;;) {
                    // two-queue merge: leaves [0, m) ascending, internal nodes [m, 2m-1) are created in ascending order
                    u32 li = 0, ii = m, ni = m;
                    for (u32 k = 0; k + 1 < m; k++) {
                        u32 a, b;
                        if (li < m && (ii >= ni || wnode[li] <= wnode[ii])) a = li++; else a = ii++;
                        if (li < m && (ii >= ni || wnode[li] <= wnode[ii])) b = li++; else b = ii++;
                        wnode[ni] = wnode[a] + wnode[b];
                        parent[a] = parent[b] = (u16)ni;
                        ni++;
                    }
                    // depths from the root down (a parent always has a larger index than its children); the weights of
                    // internal nodes are not needed any more, so their slots hold the depths
                    const u32 root = 2 * m - 2;
                    maxd = 0;
                    wnode[root] = 0;
                    for (u32 v = root; v-- > 0;) {
                        const u32 d = wnode[parent[v]] + 1u;
                        if (v >= m) wnode[v] = d;
                        else { len[order[v]] = (u8)d; if (d > maxd) maxd = d; }
                    }
                    if (maxd <= KZG_HUF_MAXBITS) break;
                    // too deep: flatten the counts and build again (the order stays valid)
                    for (u32 k = 0; k < m; k++) { const u32 c = hist[order[k]]; const u32 h2 = (c + 1) >> 1; hist[order[k]] = h2 ? h2 : 1; wnode[k] = hist[order[k]]; }
                }
                misc[3] = maxd;
            }
            __syncthreads();
            const u32 maxbits = misc[3];
            // ---- canonical codes, the way the decoder fills its table: by weight (longest codes first), then by symbol
            for (u32 t = tid; t < 256; t += blockDim.x) {
                const u32 l = len[t];
                if (l) {
                    const u32 w = maxbits + 1 - l;
                    u32 below = 0, r = 0;                                // table entries of lower weights; same-weight symbols before t
                    for (u32 j = 0; j < 256; j++) {
                        const u32 lj = len[j];
                        if (lj > l) below += 1u << (maxbits - lj);       // weight wj < w: 2^(wj-1) = 2^(maxbits-lj) entries
                        r += (lj == l) && (j < t);
                    }
                    code[t] = (u16)((below >> (w - 1)) + r);
                }
            }
            __syncthreads();
            // ---- bits of each stream
            const u32 seg = (nl + 3) >> 2;
            {
                const u32 s0 = min(warp * seg, nl), s1 = min(s0 + seg, nl);
                u32 bitsum = 0;
                for (u32 k = s0 + lane; k < s1; k += 32) bitsum += len[lit[k]];
#pragma unroll
                for (int o = 16; o; o >>= 1) bitsum += __shfl_xor_sync(0xFFFFFFFFu, bitsum, o);
                if (lane == 0) misc[4 + warp] = bitsum;
            }
            __syncthreads();
            if (tid == 0) {
                // tree description: the weights of symbols 0 .. maxsym-1 (the last one is implicit), FSE-coded when
                // that is smaller or when there are more than 128 of them, else two per byte
                for (u32 k = 0; k < maxsym; k++) wts[k] = len[k] ? (u8)(maxbits + 1 - len[k]) : (u8)0;
                const u32 direct = 1 + ((maxsym + 1) >> 1);
                const u32 fsz = kzg_enc_compress_weights(wts, maxsym, wbuf, 127, wsc);
                const bool coded = fsz != 0 && (maxsym > 128 || 1 + fsz < direct);
                const u32 whdr = coded ? 1 + fsz : direct;
                const u32 sz0 = (misc[4] + 8) >> 3, sz1 = (misc[5] + 8) >> 3, sz2 = (misc[6] + 8) >> 3, sz3 = (misc[7] + 8) >> 3;
                const u32 csize = whdr + 6 + sz0 + sz1 + sz2 + sz3;
                const u32 big = nl > csize ? nl : csize;
                const u32 fmt = big < 1024 ? 1u : (big < 16384 ? 2u : 3u), lhdr = 2 + fmt, fbits = fmt == 1 ? 10u : (fmt == 2 ? 14u : 18u);
                if (lhdr + csize >= 3 + nl || sz0 > 65535 || sz1 > 65535 || sz2 > 65535 || (!coded && maxsym > 128)) misc[2] = 0;      // raw is not larger
                else {
                    misc[2] = 2;
                    const u64 h = 2ull | ((u64)fmt << 2) | ((u64)nl << 4) | ((u64)csize << (4 + fbits));
                    for (u32 k = 0; k < lhdr; k++) blk[k] = (u8)(h >> (8 * k));
                    u8* p = blk + lhdr;
                    if (coded) {
                        *p++ = (u8)fsz;
                        for (u32 k = 0; k < fsz; k++) *p++ = wbuf[k];
                    } else {
                        *p++ = (u8)(127 + maxsym);
                        for (u32 k = 0; k < maxsym; k += 2) *p++ = (u8)((wts[k] << 4) | (k + 1 < maxsym ? wts[k + 1] : 0));
                    }
                    kzg_wr16(p, sz0); kzg_wr16(p + 2, sz1); kzg_wr16(p + 4, sz2);
                    const u32 base = lhdr + whdr + 6;
                    misc[8] = base; misc[9] = base + sz0; misc[10] = base + sz0 + sz1; misc[11] = base + sz0 + sz1 + sz2;
                    misc[12] = lhdr + csize;
                }
            }
            __syncthreads();
            mode = misc[2];
            if (mode == 2) {
                // ---- the four streams, one warp each: symbols from the last to the first, 128 per step
                const u32 s0 = min(warp * seg, nl), s1 = min(s0 + seg, nl);
                u8* out = blk + misc[8 + warp];
                u32* st = stage[warp];
                u32 carry_bits = 0, carry_n = 0;
                for (u32 hi = s1; hi > s0;) {
                    const u32 cnt = min(128u, hi - s0);
                    // this lane's (at most four) symbols, in the order they are written: hi-1-4*lane, hi-2-4*lane, ...
                    u64 v = 0;
                    u32 vn = 0;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const u32 j = 4 * lane + t;
                        if (j < cnt) {
                            const u32 sym = lit[hi - 1 - j];
                            v |= (u64)code[sym] << vn;
                            vn += len[sym];
                        }
                    }
                    const u32 incl = warp_incl_sum(vn, lane);
                    const u32 total = carry_n + __shfl_sync(0xFFFFFFFFu, incl, 31);
                    for (u32 k = lane; k < 48; k += 32) st[k] = k == 0 ? carry_bits : 0u;
                    __syncwarp();
                    if (vn) {
                        const u32 bp = carry_n + incl - vn;
                        const u32 sh = bp & 31u;
                        atomicOr(&st[bp >> 5], (u32)(v << sh));
                        const u64 hi64 = sh ? (v >> (32 - sh)) : (v >> 32);       // bits that did not fit the first word
                        if ((u32)hi64) atomicOr(&st[(bp >> 5) + 1], (u32)hi64);
                        if ((u32)(hi64 >> 32)) atomicOr(&st[(bp >> 5) + 2], (u32)(hi64 >> 32));
                    }
                    __syncwarp();
                    const u32 nbytes = total >> 3;
                    for (u32 k = lane; k < nbytes; k += 32) out[k] = (u8)(st[k >> 2] >> (8 * (k & 3)));
                    carry_n = total & 7u;
                    carry_bits = carry_n ? (st[nbytes >> 2] >> (8 * (nbytes & 3))) & 0xFFu : 0u;
                    out += nbytes;
                    hi -= cnt;
                    __syncwarp();
                }
                if (lane == 0) *out = (u8)(carry_bits | (1u << carry_n));       // end mark above the last bits
                if (tid == 0) B.chunks[i].lit_section = misc[12];
            }
        }
        if (mode == 1) {
            if (tid == 0) { kzg_enc_write_lit_header(blk, 1, nl); blk[3] = lit[0]; B.chunks[i].lit_section = 4; }
        } else if (mode == 0) {
            if (tid == 0) { kzg_enc_write_lit_header(blk, 0, nl); B.chunks[i].lit_section = 3 + nl; }
            for (u32 k = tid; k < nl; k += blockDim.x) blk[3 + k] = lit[k];
        }
    }
}

// encode stage: one warp per chunk.  Shared memory per warp: the three FSE tables' lookups go to global memory
// (constant, cached); 3 x 32 words of chain results and a 640-byte bit staging area live here.
#define KZG_ENC_EWARPS 4
#define KZG_ENC_STAGE_WORDS 96       // 32 sequences x at most 66 bits = 2112 bits < 96 words, plus the carried partial byte
__global__ void __launch_bounds__(KZG_ENC_EWARPS * 32) kzg_enc_encode_kernel(KzgEncBatch B) {
    using namespace kzg_encd;
    __shared__ u32 s_chain[KZG_ENC_EWARPS][3][32];
    __shared__ u32 s_bits[KZG_ENC_EWARPS][KZG_ENC_STAGE_WORDS];
    __shared__ KzgEncTables s_tab;
    __shared__ KzgEncTabSet s_set[KZG_ENC_EWARPS];         // the tables the warp's current block is written with
    __shared__ KzgEncTabScratch s_scr[KZG_ENC_EWARPS];
    __shared__ u32 s_hist[KZG_ENC_EWARPS][128];            // code histograms: LL at 0, ML at 36, OF at 89
    __shared__ u32 s_syms[KZG_ENC_EWARPS][100];            // per warp: codes of the batch (3 x 32), final states (3)
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 k = threadIdx.x; k < sizeof(KzgEncTables) / 4; k += blockDim.x) reinterpret_cast<u32*>(&s_tab)[k] = reinterpret_cast<const u32*>(B.tab)[k];
    __syncthreads();
    KzgEncTabSet& ts = s_set[warp];
    u32* hist = s_hist[warp];
    u32* syms = s_syms[warp];
    const KzgEncTables* T = &s_tab;
    u32* chain = &s_chain[warp][0][0];
    u32* bits = s_bits[warp];
    for (;;) {
        u32 i = 0;
        if (lane == 0) i = atomicAdd(&B.tickets[2], 1u);
        i = __shfl_sync(0xFFFFFFFFu, i, 0);
        if (i >= B.n) break;
        const KzgEncChunk ck = B.chunks[i];
        u32 nseq = 0;
#pragma unroll
        for (int q = 0; q < KZG_ENC_WARPS; q++) nseq += ck.n_seq[q];
        const u32 qcap = B.seq_cap / KZG_ENC_WARPS;
        u8 *blk = nullptr, *limit = nullptr, *p = nullptr;
        if (lane == 0) p = kzg_enc_block_prologue(B, i, nseq, blk, limit);
        p = reinterpret_cast<u8*>(__shfl_sync(0xFFFFFFFFu, (unsigned long long)p, 0));
        if (!p) continue;
        blk = reinterpret_cast<u8*>(__shfl_sync(0xFFFFFFFFu, (unsigned long long)blk, 0));
        limit = reinterpret_cast<u8*>(__shfl_sync(0xFFFFFFFFu, (unsigned long long)limit, 0));
        u64* seq = B.seq + (size_t)i * B.seq_cap;
        // ---- forward: offset values with repeat codes.  Loads and stores are coalesced, the history walk runs on
        // every lane redundantly (uniform control flow, no divergence) over shuffled records.
        for (u32 k = lane; k < 128; k += 32) hist[k] = 0;
        __syncwarp();
        {
            u32 r0 = 1, r1 = 4, r2 = 8;
#pragma unroll 1
            for (int q = 0; q < KZG_ENC_WARPS; q++) {
                const u32 nq = B.chunks[i].n_seq[q];
                u64* sq = seq + (size_t)q * qcap;
                for (u32 b = 0; b < nq; b += 32) {
                    const bool have = b + lane < nq;
                    const u64 rec = have ? sq[b + lane] : 0ull;
                    const u32 cnt = min(32u, nq - b);
                    u32 mine = 0;
                    for (u32 j = 0; j < cnt; j++) {
                        const u64 rj = __shfl_sync(0xFFFFFFFFu, rec, j);
                        const u32 ob = kzg_enc_offset_value((u32)rj & 0x3FFFFu, (u32)(rj >> 36) & 0x3FFFFu, r0, r1, r2);
                        if (j == lane) mine = ob;
                    }
                    if (have) {
                        sq[b + lane] = (rec & 0xFFFFFFFFFull) | ((u64)mine << 36);
                        atomicAdd(&hist[kzg_enc_ll_code(T, (u32)rec & 0x3FFFFu)], 1u);
                        atomicAdd(&hist[36 + kzg_enc_ml_code(T, ((u32)(rec >> 18) & 0x3FFFFu) - 3u)], 1u);
                        atomicAdd(&hist[89 + (u32)kzg_highbit(mine)], 1u);
                    }
                }
            }
        }
        __syncwarp();
        // ---- the tables of this block: modes, descriptions, encoding tables (serial, on one lane)
        if (lane == 0) p = kzg_enc_prepare_tables(T, nseq, hist, ts, s_scr[warp], p, limit);
        p = reinterpret_cast<u8*>(__shfl_sync(0xFFFFFFFFu, (unsigned long long)p, 0));
        if (!p) continue;                                        // no room: stays a raw block
        __syncwarp();
        // ---- backward: batches of at most 32 sequences inside one part, lane j <-> sequence hi - 1 - j (the order
        // they are written in)
        u32 stLL = 0, stML = 0, stOF = 0;          // current states (kept on every lane)
        u32 carry_bits = 0, carry_n = 0;            // bits of the last, incomplete byte
        u8* out = p;
        bool ovf = false, first = true;
#pragma unroll 1
        for (int q = KZG_ENC_WARPS - 1; q >= 0 && !ovf; q--) {
            const u64* sq = seq + (size_t)q * qcap;
            for (u32 hi = B.chunks[i].n_seq[q]; hi > 0;) {
                const u32 cnt = min(32u, hi);
                const bool have = lane < cnt;
                const u64 rec = have ? sq[hi - 1 - lane] : 0ull;
                const u32 ll = (u32)rec & 0x3FFFFu, mlb = have ? ((u32)(rec >> 18) & 0x3FFFFu) - 3u : 0u, ob = have ? (u32)(rec >> 36) & 0x3FFFFu : 1u;
                const u32 cl = kzg_enc_ll_code(T, ll), cm = kzg_enc_ml_code(T, mlb), co = (u32)kzg_highbit(ob);
                // the three state chains on lanes 0 (LL), 1 (ML), 2 (OF) over the codes parked in shared memory;
                // results: value bits | nb << 16 per sequence
                syms[lane] = cl; syms[32 + lane] = cm; syms[64 + lane] = co;
                __syncwarp();
                if (lane < 3) {
                    const u16* st = ts.st + KZG_ENC_ST_OFF(lane);
                    const u32* dnb = ts.dnb + KZG_ENC_SYM_OFF(lane);
                    const i32* dfs = ts.dfs + KZG_ENC_SYM_OFF(lane);
                    const u32* my = syms + 32 * lane;
                    u32* res = chain + 32 * lane;
                    u32 value = lane == 0 ? stLL : (lane == 1 ? stML : stOF);
                    u32 j = 0;
                    if (first) {                               // the very last sequence initialises the states
                        const u32 sym = my[0];
                        if (ts.log[lane] == 0) value = 1;      // RLE mode: the state carries no bits
                        else {
                            const u32 nb = (dnb[sym] + (1u << 15)) >> 16;
                            value = st[(i32)(((nb << 16) - dnb[sym]) >> nb) + dfs[sym]];
                        }
                        res[0] = 0;
                        j = 1;
                    }
                    for (; j < cnt; j++) {
                        const u32 sym = my[j];
                        const u32 nb = (value + dnb[sym]) >> 16;
                        res[j] = (value & ((1u << nb) - 1u)) | (nb << 16);
                        value = st[(i32)(value >> nb) + dfs[sym]];
                    }
                    syms[96 + lane] = value;
                }
                __syncwarp();
                stLL = syms[96]; stML = syms[97]; stOF = syms[98];
                first = false;
                __syncwarp();
                // this lane's fields, in stream order: OF state, ML state, LL state, LL extra, ML extra, OF extra
                u32 fv[6], fn[6];
                {
                    const u32 eL = chain[lane], eM = chain[32 + lane], eO = chain[64 + lane];
                    fv[0] = eO & 0xFFFFu; fn[0] = eO >> 16;
                    fv[1] = eM & 0xFFFFu; fn[1] = eM >> 16;
                    fv[2] = eL & 0xFFFFu; fn[2] = eL >> 16;
                    fn[3] = T->ll_tab[cl] >> 24; fv[3] = ll & ((1u << fn[3]) - 1u);
                    fn[4] = T->ml_tab[cm] >> 24; fv[4] = mlb & ((1u << fn[4]) - 1u);
                    fn[5] = co; fv[5] = ob & ((1u << co) - 1u);
                    if (!have) { fn[0] = fn[1] = fn[2] = fn[3] = fn[4] = fn[5] = 0; }
                }
                const u32 mylen = fn[0] + fn[1] + fn[2] + fn[3] + fn[4] + fn[5];
                const u32 incl = warp_incl_sum(mylen, lane);
                const u32 total = carry_n + __shfl_sync(0xFFFFFFFFu, incl, 31);
                for (u32 k = lane; k < KZG_ENC_STAGE_WORDS; k += 32) bits[k] = k == 0 ? carry_bits : 0u;
                __syncwarp();
                u32 bp = carry_n + incl - mylen;
#pragma unroll
                for (int f = 0; f < 6; f++) {
                    if (fn[f]) {
                        const u64 v = (u64)fv[f] << (bp & 31u);
                        atomicOr(&bits[bp >> 5], (u32)v);
                        if ((u32)(v >> 32)) atomicOr(&bits[(bp >> 5) + 1], (u32)(v >> 32));
                    }
                    bp += fn[f];
                }
                __syncwarp();
                // whole bytes leave; the incomplete last byte is carried into the next batch
                const u32 nbytes = total >> 3;
                if (out + nbytes > limit) { ovf = true; break; }
                for (u32 k = lane; k < nbytes; k += 32) out[k] = (u8)(bits[k >> 2] >> (8 * (k & 3)));
                carry_n = total & 7u;
                carry_bits = carry_n ? (bits[nbytes >> 2] >> (8 * (nbytes & 3))) & 0xFFu : 0u;
                out += nbytes;
                hi -= cnt;
                __syncwarp();
            }
        }
        if (ovf) continue;                                       // not smaller than raw: stays ST_RAW
        if (lane == 0) {
            KzgBitW w;
            w.acc = carry_bits; w.nbits = carry_n; w.p = out; w.end = limit; w.ovf = false;
            kzg_bw_add(w, stML, ts.log[1]);
            kzg_bw_add(w, stOF, ts.log[2]);
            kzg_bw_add(w, stLL, ts.log[0]);
            kzg_bw_add(w, 1, 1);                                 // end mark
            kzg_bw_flush(w);
            if (w.nbits) { if (w.p < w.end) *w.p = (u8)w.acc; else w.ovf = true; w.p++; }
            if (!w.ovf && w.p < limit) kzg_enc_block_epilogue(B, i, blk, w.p);
        }
        __syncwarp();
    }
}

// finish stage: raw-block fallback, checksum trailer, sizes.  One warp per chunk.
__global__ void __launch_bounds__(128) kzg_enc_finish_kernel(KzgEncBatch B) {
    using namespace kzg_encd;
    const u32 lane = threadIdx.x & 31;
    for (;;) {
        u32 i = 0;
        if (lane == 0) i = atomicAdd(&B.tickets[1], 1u);
        i = __shfl_sync(0xFFFFFFFFu, i, 0);
        if (i >= B.n) break;
        const KzgEncChunk ck = B.chunks[i];
        if (ck.state == KZG_ENC_ST_ERROR) continue;
        const u32 n = (u32)B.src_size[i];
        const u8* s = B.src + B.src_off[i];
        u8* dst = B.dst + B.dst_off[i];
        u32 total = ck.total;
        if (ck.state != KZG_ENC_ST_COMPRESSED) {
            const u32 hdr = kzg_enc_header_size(n);
            if (lane == 0) kzg_wr24(dst + hdr, 1u | (n << 3));          // last block, raw
            warp_copy_to_end(dst + hdr + 3, s, n, lane);
            total = hdr + 3 + n;
        }
        if (B.flags & KZG_ENC_FLAG_CHECKSUM) {
            // XXH64 (Appendix A.8): lanes 0..3 are the four accumulators
            const u64 P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full, P3 = 0x165667B19E3779F9ull,
                      P4 = 0x85EBCA77C2B2AE63ull, P5 = 0x27D4EB2F165667C5ull;
            auto rotl = [](u64 x, int r) { return (x << r) | (x >> (64 - r)); };
            auto round = [&](u64 a, u64 x) { return rotl(a + x * P2, 31) * P1; };
            u64 acc = lane == 0 ? P1 + P2 : (lane == 1 ? P2 : (lane == 2 ? 0ull : 0ull - P1));
            const u32 stripes = n >> 5;
            if (lane < 4) for (u32 k = 0; k < stripes; k++) acc = round(acc, kzg_rd64(s + 32 * k + 8 * lane));
            const u64 v2 = __shfl_sync(0xFFFFFFFFu, acc, 1), v3 = __shfl_sync(0xFFFFFFFFu, acc, 2), v4 = __shfl_sync(0xFFFFFFFFu, acc, 3);
            if (lane == 0) {
                u64 h;
                if (n >= 32) {
                    h = rotl(acc, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
                    h = (h ^ round(0, acc)) * P1 + P4; h = (h ^ round(0, v2)) * P1 + P4;
                    h = (h ^ round(0, v3)) * P1 + P4; h = (h ^ round(0, v4)) * P1 + P4;
                } else h = P5;
                h += n;
                const u8* c = s + (n & ~31u);
                const u8* e = s + n;
                while (e - c >= 8) { h = rotl(h ^ round(0, kzg_rd64(c)), 27) * P1 + P4; c += 8; }
                if (e - c >= 4) { h = rotl(h ^ ((u64)kzg_rd32(c) * P1), 23) * P2 + P3; c += 4; }
                while (c < e) { h = rotl(h ^ ((u64)(*c) * P5), 11) * P1; c++; }
                h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
                kzg_wr32(dst + total, (u32)h);
            }
            total += 4;
        }
        if (lane == 0) { B.status[i] = KZG_OK; B.out_size[i] = total; }
    }
}
#endif  // __CUDACC__
