// Entropy stage bodies: FSE decode-table build, Huffman decode-table build, Huffman stream decode
// (one thread per stream) and the ZSTD_decodeSequence state machine (one thread per block).
// Serial per-thread code (KZG_HD) so that tests/hostsim can run the identical logic on a CPU.
#pragma once
#include "kzg_bits.cuh"

// ---------------------------------------------------------------------------------------------
// FSE decode table, 2 bytes per cell: symbol in bits [0,6), `next` in bits [6,16).
// nbBits and baseline are recomputed at decode time from `next` (Appendix A.5):
//   nbBits = log - floor(log2(next)),  baseline = (next << nbBits) - (1 << log).
// 2-byte cells halve the shared-memory footprint (2560 B per block for logs 9/8/9), which is what
// bounds how many blocks one SM decodes concurrently.
// 16-byte copy (global -> shared on the device); both pointers 16-byte aligned.
KZG_HD void kzg_copy16(void* dst, const void* src) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
#else
    memcpy(dst, src, 16);
#endif
}

// `scratch` (16-byte aligned, >= 2*ceil8(nsym) bytes, thread-private shared memory) first receives the
// normalised counts with wide loads, then serves as the per-symbol `next` counters: no per-element
// global or local-memory round trip sits on the build's dependency chain.
// scratch encoding: count, or 0x8001 for a "less than one" (-1) symbol.
KZG_HD void kzg_fse_build(u16* tab, const i16* norm, u32 nsym, u32 log, u16* scratch) {
    const u32 size = 1u << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    u32 high = size, pos = 0;
    const u32 nchunks = (nsym + 7) >> 3;
    for (u32 k = 0; k < nchunks; k++) kzg_copy16(scratch + 8 * k, norm + 8 * k);
    for (u32 s = 0; s < nsym; s++) {
        if (scratch[s] == 0xFFFFu) { tab[--high] = (u16)s; scratch[s] = 0x8001u; }
    }
    for (u32 s = 0; s < nsym; s++) {
        const u32 v = scratch[s];
        const u32 c = (v & 0x8000u) ? 0u : v;
        for (u32 i = 0; i < c; i++) {
            tab[pos] = (u16)s;
            do { pos = (pos + step) & mask; } while (pos >= high);
        }
    }
    for (u32 i = 0; i < size; i++) {
        const u32 s = tab[i];
        const u32 v = scratch[s];
        scratch[s] = (u16)(v + 1);
        tab[i] = (u16)(((v & 0x7FFFu) << 6) | s);
    }
}

// Builds table `t` (0 LL, 1 OF, 2 ML) described by `d` into tab; returns its accuracy log.
KZG_HD u32 kzg_fse_build_desc(u16* tab, const KzgFseDesc* d, u32 t, u16* scratch) {
    if (d->kind[t] == 1) { tab[0] = (u16)((1u << 6) | (u32)d->norm[t][0]); return 0; }
    kzg_fse_build(tab, d->norm[t], d->nsym[t], d->log[t], scratch);
    return d->log[t];
}

// Predefined distributions (Appendix A.6) as a descriptor; the API uploads one copy to the device.
inline void kzg_default_fse_desc(KzgFseDesc& d) {
    static const i16 LL[36] = { 4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1 };
    static const i16 OF[29] = { 1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1 };
    static const i16 ML[53] = { 1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1 };
    for (int i = 0; i < 64; i++) { d.norm[0][i] = i < 36 ? LL[i] : 0; d.norm[1][i] = i < 29 ? OF[i] : 0; d.norm[2][i] = i < 53 ? ML[i] : 0; }
    d.log[0] = 6; d.log[1] = 5; d.log[2] = 6;
    d.nsym[0] = 36; d.nsym[1] = 29; d.nsym[2] = 53;
    d.kind[0] = d.kind[1] = d.kind[2] = 2;
    for (int i = 0; i < 7; i++) d.pad[i] = 0;
}

// Literal-length / match-length code tables packed as (extra_bits << 24) | baseline (Appendix A.6).
inline void kzg_code_tables(u32* ll36, u32* ml53) {
    static const u32 LLB[36] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,18,20,22,24,28,32,40,48,64,128,256,512,1024,2048,4096,8192,16384,32768,65536 };
    static const u8 LLX[36] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16 };
    static const u32 MLB[53] = { 3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32,33,34,35,37,39,41,43,47,51,59,67,83,99,131,259,515,1027,2051,4099,8195,16387,32771,65539 };
    static const u8 MLX[53] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16 };
    for (int i = 0; i < 36; i++) ll36[i] = ((u32)LLX[i] << 24) | LLB[i];
    for (int i = 0; i < 53; i++) ml53[i] = ((u32)MLX[i] << 24) | MLB[i];
}

// ---------------------------------------------------------------------------------------------
// Huffman decode table (Appendix A.4), 2 bytes per entry: symbol | code_length << 8.
// The four threads that will decode a block's four streams build its table together. Each lane owns
// the 64 symbols [64*lane, 64*lane+64): it stages their weights with wide loads (phase 1), counts them,
// derives where its symbols start inside every weight class (phase 2) and fills their spans (phase 3).
// Spans are powers of two and naturally aligned. The caller synchronises the lanes between phases.
struct KzgHufScratch { u8 w[256]; u16 cnt[4][16]; u16 st[4][16]; };   // 512 B per table, shared memory

KZG_HD void kzg_huf_build_phase1(KzgHufScratch* sc, const u8* weights, u32 nsym, u32 lane) {
    for (u32 k = 0; k < 4; k++) kzg_copy16(sc->w + 64 * lane + 16 * k, weights + 64 * lane + 16 * k);
    for (u32 w = 0; w < 16; w++) sc->cnt[lane][w] = 0;
    const u32 lo = 64 * lane, hi = nsym < lo + 64 ? nsym : lo + 64;
    for (u32 s = lo; s < hi; s++) sc->cnt[lane][sc->w[s] & 15]++;
}

KZG_HD void kzg_huf_build_phase2(KzgHufScratch* sc, u32 maxbits, u32 lane) {
    u32 pos = 0;
    for (u32 w = 1; w <= maxbits; w++) {
        u32 before = 0, tot = 0;
        for (u32 l = 0; l < 4; l++) { const u32 c = sc->cnt[l][w]; if (l < lane) before += c; tot += c; }
        sc->st[lane][w] = (u16)(pos + (before << (w - 1)));
        pos += tot << (w - 1);
    }
}

// Two-level table.  The full table would have 2^maxbits entries (4 KiB at 11 bits), which limits how many blocks an
// SM can decode at once.  The canonical order puts the longest codes first, so the entries of the codes longer than
// KZG_HUF_L1_BITS are exactly the first X entries of the full table (X = sum over the weights w <= sshift of
// count_w * 2^(w-1), a multiple of 2^sshift by the Kraft equality) and every other code covers whole groups of 2^sshift
// entries.  Stored are therefore
//   l2[0, X)                       = the first X entries of the full table (long codes), and
//   l1[0, 2^(maxbits - sshift))    = one entry per group of 2^sshift (short codes; the groups below X >> sshift unused),
// and a lookup is ONE load: index idx (the next maxbits bits) reads l2[idx] if idx < X, else l1[idx >> sshift].
// At most 1 KiB + 2 * X bytes per table; on text X is a few dozen.
// With l1bits = KZG_HUF_MAXBITS the first level is the full table, X = 0 and the second level is empty: the layout of
// the launch that takes the blocks with long streams (there the chain latency per symbol counts, not the footprint).
#define KZG_HUF_L1_BITS 9
KZG_HD u32 kzg_huf_split(u32 maxbits, u32 l1bits = KZG_HUF_L1_BITS) { return maxbits > l1bits ? maxbits - l1bits : 0u; }

// (after phase 1) entries of the full table that belong to codes longer than l1bits
KZG_HD u32 kzg_huf_long_entries(const KzgHufScratch* sc, u32 maxbits, u32 l1bits = KZG_HUF_L1_BITS) {
    const u32 sshift = kzg_huf_split(maxbits, l1bits);
    u32 x = 0;
    for (u32 w = 1; w <= sshift; w++) x += ((u32)sc->cnt[0][w] + sc->cnt[1][w] + sc->cnt[2][w] + sc->cnt[3][w]) << (w - 1);
    return x;
}

KZG_HD void kzg_huf_build_phase3(u16* l1, u16* l2, KzgHufScratch* sc, u32 nsym, u32 maxbits, u32 lane, u32 l1bits = KZG_HUF_L1_BITS) {
    const u32 lo = 64 * lane, hi = nsym < lo + 64 ? nsym : lo + 64;
    const u32 sshift = kzg_huf_split(maxbits, l1bits);
    for (u32 s = lo; s < hi; s++) {
        const u32 w = sc->w[s];
        if (w == 0 || w > maxbits) continue;
        const u32 span = 1u << (w - 1), p = sc->st[lane][w];
        sc->st[lane][w] = (u16)(p + span);
        const u32 e = s | ((maxbits + 1 - w) << 8);
        if (w <= sshift) {                       // a long code: 1 or 2 entries of the second level
            l2[p] = (u16)e;
            if (span == 2) l2[p + 1] = (u16)e;
            continue;
        }
        const u32 q = p >> sshift, n = span >> sshift;       // naturally aligned run of n = 2^k first-level entries
        if (n == 1) { l1[q] = (u16)e; continue; }
        const u32 e2 = e | (e << 16);
        u32* t32 = reinterpret_cast<u32*>(l1) + (q >> 1);
        for (u32 k = 0; k < (n >> 1); k++) t32[k] = e2;
    }
}

// One Huffman stream: decodes exactly n symbols from [src, src+size) into dst. Returns 0 or an error.
KZG_HD u32 kzg_huf_decode_stream(const u8* src, u32 size, const u16* l1, const u16* l2, u32 nlong, u32 maxbits, u8* dst, u32 n) {
    KzgBwdQ b;
    if (!b.init(src, size)) return KZG_E_CORRUPTION;
    const u32 shift = 64 - maxbits, sshift = kzg_huf_split(maxbits);
    u32 i = 0;
#define KZG_HUF_SYM(out)                                   \
    {                                                      \
        const u32 idx = (u32)(b.win >> shift);             \
        const u32 e = idx < nlong ? l2[idx] : l1[idx >> sshift]; \
        b.skip(e >> 8);                                    \
        out = e & 255u;                                    \
    }
    // head: bytes until dst is 16-byte aligned
    while (i < n && (((uintptr_t)(dst + i)) & 15)) {
        b.ensure32();
        u32 s; KZG_HUF_SYM(s);
        dst[i++] = (u8)s;
    }
    // body: 16 symbols per 16-byte store
    while (i + 16 <= n) {
        b.advance();                              // warp-uniform point (every 16 symbols)
        u32 v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            u32 s0, s1, s2, s3;
            b.ensure32(); KZG_HUF_SYM(s0); KZG_HUF_SYM(s1);
            b.ensure32(); KZG_HUF_SYM(s2); KZG_HUF_SYM(s3);
            v[q] = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
        }
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<uint4*>(dst + i) = make_uint4(v[0], v[1], v[2], v[3]);
#else
        memcpy(dst + i, v, 16);
#endif
        i += 16;
    }
    b.advance();
    while (i < n) {
        b.ensure32();
        u32 s; KZG_HUF_SYM(s);
        dst[i++] = (u8)s;
    }
#undef KZG_HUF_SYM
    return b.left == 0 ? 0u : (b.left > 0 ? 1u : 2u);       // exactly consumed / bits left over / ran out of bits
}

// What a Huffman stream that was not consumed exactly means.  RFC 8878 4.2.2 wants every stream consumed exactly, and
// libzstd 1.5.7 checks that in its plain loops (one stream, or four streams one of which is shorter than 8 bytes).  Its
// fast four-stream loops (HUF_decompress4X*_usingDTable_internal_fast) only verify the symbol count: left-over bits are
// accepted (same bytes here), and a stream that runs out of bits yields literals made up from a stale bit container,
// which this decoder does not reproduce -- the block is flagged and its frame ends as checksum_wrong when it carries a
// checksum (libzstd's verdict, a 2^-32 chance aside) and as corruption_detected when it does not.
//   ends: 0 exact, 1 bits left over, 2 ran out of bits.  Returns true if the block is corrupt; may set cb->lit_garbage.
KZG_HD bool kzg_huf_verdict(KzgCBlock* cb, u32 ends) {
    if (ends == 0) return false;
    const u32 m = cb->n_streams == 4 ? (cb->stream_size[0] < 8u || cb->stream_size[1] < 8u || cb->stream_size[2] < 8u || cb->stream_size[3] < 8u) : 1u;
    if (m) return true;
    if (ends == 2) cb->lit_garbage = 1u;
    return false;
}

// ---------------------------------------------------------------------------------------------
// Repeat-offset arithmetic on encoded values (see KZG_OFF_SYM in kzg_common.cuh).
KZG_HD u32 kzg_rep_dec(u32 r) {          // "rep0 - 1"
    if (r & KZG_OFF_SYM) return r + 4;   // symbolic: d += 1
    if (r == KZG_OFF_INVALID || r <= 1) return KZG_OFF_INVALID;
    return r - 1;
}

struct KzgSeqOut { u32 status, regen, lit_used, rep0, rep1, rep2, bulk_lits; };

// ZSTD_decodeSequence loop for one block (Appendix A.6): the sequential state machine.
//  bits/size: sequence bitstream; tLL/tOF/tML: decode tables; llx/mlx: packed code tables;
//  rep*: encoded repeat offsets at block start; out: nseq packed records.
KZG_HD void kzg_decode_sequences(const u8* bits, u32 size, u32 nseq,
                                 const u16* tLL, u32 logLL, const u16* tOF, u32 logOF, const u16* tML, u32 logML,
                                 const u32* llx, const u32* mlx, u32 rep0, u32 rep1, u32 rep2,
                                 u32 lit_regen, u32 block_max, u64* out, KzgSeqOut& res) {
    KzgBwdQ b;
    res.status = KZG_E_CORRUPTION; res.regen = 0; res.lit_used = 0; res.bulk_lits = 0;
    res.rep0 = rep0; res.rep1 = rep1; res.rep2 = rep2;
    if (!b.init(bits, size)) return;
    b.ensure32();
    u32 sLL = b.read(logLL), sOF = b.read(logOF), sML = b.read(logML);
    if (b.left < 0) return;
    u32 lit_end = 0, out_end = 0;
    const u32 szLL = 1u << logLL, szOF = 1u << logOF, szML = 1u << logML;
    for (u32 i = 0; i < nseq; i++) {
        if ((i & 7u) == 7u) b.advance();          // warp-uniform point: the only place global loads are issued
        const u32 eLL = tLL[sLL], eOF = tOF[sOF], eML = tML[sML];
        const u32 ofc = eOF & 63u;
        const u32 pLL = llx[eLL & 63u], pML = mlx[eML & 63u];
        b.ensure32();
        const u32 ofv = (1u << ofc) + b.read(ofc);
        b.ensure32();
        const u32 ml = (pML & 0xFFFFFFu) + b.read(pML >> 24);
        const u32 ll = (pLL & 0xFFFFFFu) + b.read(pLL >> 24);
        u32 off;
        if (ofv > 3) {
            off = ofv - 3;
            if (off >= KZG_OFF_INVALID) off = KZG_OFF_INVALID;
            rep2 = rep1; rep1 = rep0; rep0 = off;
        } else {
            const u32 idx = ofv - 1 + (ll == 0);
            if (idx == 0) off = rep0;
            else {
                off = idx == 3 ? kzg_rep_dec(rep0) : (idx == 1 ? rep1 : rep2);
                if (idx != 1) rep2 = rep1;
                rep1 = rep0; rep0 = off;
            }
        }
        lit_end += ll;
        out_end += ll + ml;
        if (ll >= KZG_BULK_MIN) res.bulk_lits += ll;          // sum of long runs; turned into the flag by the caller
        if (lit_end > lit_regen || out_end > block_max) return;
        out[i] = kzg_pack_seq(off, lit_end, out_end);
        if (i + 1 < nseq) {
            b.ensure32();
            const u32 nLL = eLL >> 6, nML = eML >> 6, nOF = eOF >> 6;
            const u32 bLL = logLL - (u32)kzg_highbit(nLL), bML = logML - (u32)kzg_highbit(nML), bOF = logOF - (u32)kzg_highbit(nOF);
            sLL = (nLL << bLL) - szLL + b.read(bLL);
            sML = (nML << bML) - szML + b.read(bML);
            sOF = (nOF << bOF) - szOF + b.read(bOF);
        }
    }
    if (b.left != 0) return;
    const u32 regen = out_end + (lit_regen - lit_end);
    if (regen > block_max) return;
    res.status = KZG_OK; res.regen = regen; res.lit_used = lit_end;
    res.rep0 = rep0; res.rep1 = rep1; res.rep2 = rep2;
}
