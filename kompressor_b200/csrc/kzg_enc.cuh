// Batch frame writer for sm_100a ("the step before the path"): compresses many
// independent chunks (<= 128 KiB each, 64 KiB by default) into one single-block Zstandard frame each, so that
// producers emit the batch shape the decoder is built for.  What the reference does here is
// ZSTD_compressStream2 (kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:75-121); any conforming
// encoder is interchangeable with it, because decoders only see the format (RFC 8878).
// Parity bar for this side: libzstd 1.5.7 decodes every emitted frame to the original bytes.
//
// Format choices of this first version: one compressed block per frame with raw literals and the predefined
// FSE tables (Appendix A.6), repeat-offset codes, optional XXH64 trailer; a raw block when that is not smaller.
//   match    warp per chunk: 32 positions per step probe a 4-byte hash table in shared memory, the first hit
//            is extended by the whole warp (128 bytes per step), literals go straight to their final place
//   encode   thread per chunk: offsets -> repeat codes (forward), then the backward FSE bitstream and headers
//   finish   warp per chunk: raw-block fallback copy, XXH64 trailer, sizes
// kzg_enc_stage_encode is KZG_HD: tests/hostsim runs the very same code on the CPU behind a sequential matcher.
#pragma once
#include "kzg_stages.cuh"

#define KZG_ENC_HASH_LOG 12
#define KZG_ENC_WARPS 4
#define KZG_ENC_FLAG_CHECKSUM 1u
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

struct KzgEncChunk { u32 n_seq, lit_size, state, total; };

struct KzgEncBatch {
    const u8* src; const u64* src_off; const u64* src_size;
    u8* dst; const u64* dst_off; const u64* dst_cap;
    u32* status; u64* out_size;
    u32 n, flags;
    u64* seq; u32 seq_cap;                            // seq_cap records per chunk: ll[0,18) | ml[18,36) | off[36,54)
    KzgEncChunk* chunks;
    const KzgEncTables* tab;
    u32* tickets;                                     // [0] match, [1] finish
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
inline void kzg_enc_build_ctable(const i16* norm, u32 nsym, u32 log, u16* state_table, u32* dnb, i32* dfs) {
    const u32 size = 1u << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    u32 cumul[66];
    u8 symbol_of[512];
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
    kzg_enc_build_ctable(d.norm[0], 36, 6, t.st_ll, t.dnb_ll, t.dfs_ll);
    kzg_enc_build_ctable(d.norm[1], 29, 5, t.st_of, t.dnb_of, t.dfs_of);
    kzg_enc_build_ctable(d.norm[2], 53, 6, t.st_ml, t.dnb_ml, t.dfs_ml);
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

// encode stage, one thread per chunk.  The match stage has left: literals at dst + header + 6, records, counts.
KZG_HD void kzg_enc_stage_encode(const KzgEncBatch& B, u32 i) {
    const u64 n = B.src_size[i];
    u8* const dst = B.dst + B.dst_off[i];
    KzgEncChunk& ck = B.chunks[i];
    const bool checksum = (B.flags & KZG_ENC_FLAG_CHECKSUM) != 0;
    ck.total = 0;
    if (n > KZG_BLOCK_MAX) { ck.state = KZG_ENC_ST_ERROR; B.status[i] = KZG_E_SRCSIZE_WRONG; B.out_size[i] = 0; return; }
    if (B.dst_cap[i] < kzg_enc_bound(n)) { ck.state = KZG_ENC_ST_ERROR; B.status[i] = KZG_E_DSTSIZE_TOO_SMALL; B.out_size[i] = 0; return; }
    const u32 hdr = kzg_enc_write_header(dst, n, checksum);
    const u32 nseq = ck.n_seq, lit_size = ck.lit_size;
    ck.state = KZG_ENC_ST_RAW;
    if (nseq == 0) return;
    const KzgEncTables* T = B.tab;
    u8* const blk = dst + hdr + 3;
    blk[0] = (u8)((3u << 2) | ((lit_size & 15u) << 4));          // raw literals, 20-bit size (3-byte header)
    blk[1] = (u8)(lit_size >> 4);
    blk[2] = (u8)(lit_size >> 12);
    u8* p = blk + 3 + lit_size;
    u8* const limit = blk + n;                                   // a compressed block must be smaller than the raw one
    if (p + 4 >= limit) return;
    if (nseq < 128) *p++ = (u8)nseq;
    else if (nseq < 0x7F00) { *p++ = (u8)((nseq >> 8) + 0x80); *p++ = (u8)nseq; }
    else { *p++ = 0xFF; kzg_wr16(p, nseq - 0x7F00); p += 2; }
    *p++ = 0;                                                    // predefined LL / OF / ML tables
    u64* seq = B.seq + (size_t)i * B.seq_cap;
    // forward: distances -> offset values with repeat codes (Appendix A.7), in place
    {
        u32 r0 = 1, r1 = 4, r2 = 8;
        for (u32 s = 0; s < nseq; s++) {
            const u64 rec = seq[s];
            const u32 ll = (u32)rec & 0x3FFFFu, off = (u32)(rec >> 36) & 0x3FFFFu;
            u32 ob;
            if (ll) {
                if (off == r0) ob = 1;
                else if (off == r1) { ob = 2; r1 = r0; r0 = off; }
                else if (off == r2) { ob = 3; r2 = r1; r1 = r0; r0 = off; }
                else { ob = off + 3; r2 = r1; r1 = r0; r0 = off; }
            } else {
                if (off == r1) { ob = 1; r1 = r0; r0 = off; }
                else if (off == r2) { ob = 2; r2 = r1; r1 = r0; r0 = off; }
                else if (off == r0 - 1 && off) { ob = 3; r2 = r1; r1 = r0; r0 = off; }
                else { ob = off + 3; r2 = r1; r1 = r0; r0 = off; }
            }
            seq[s] = (rec & 0xFFFFFFFFFull) | ((u64)ob << 36);
        }
    }
    // backward: the last sequence initialises the states, the first is written last
    KzgBitW w;
    w.acc = 0; w.nbits = 0; w.p = p; w.end = limit; w.ovf = false;
    KzgFseC sLL, sML, sOF;
    {
        const u64 rec = seq[nseq - 1];
        const u32 ll = (u32)rec & 0x3FFFFu, mlb = ((u32)(rec >> 18) & 0x3FFFFu) - 3u, ob = (u32)(rec >> 36) & 0x3FFFFu;
        const u32 cl = kzg_enc_ll_code(T, ll), cm = kzg_enc_ml_code(T, mlb), co = (u32)kzg_highbit(ob);
        kzg_fsec_init(sML, T->st_ml, T->dnb_ml, T->dfs_ml, cm);
        kzg_fsec_init(sOF, T->st_of, T->dnb_of, T->dfs_of, co);
        kzg_fsec_init(sLL, T->st_ll, T->dnb_ll, T->dfs_ll, cl);
        kzg_bw_add(w, ll, T->ll_tab[cl] >> 24);
        kzg_bw_add(w, mlb, T->ml_tab[cm] >> 24);
        kzg_bw_flush(w);
        kzg_bw_add(w, ob, co);
        kzg_bw_flush(w);
    }
    for (u32 s = nseq - 1; s-- > 0;) {
        const u64 rec = seq[s];
        const u32 ll = (u32)rec & 0x3FFFFu, mlb = ((u32)(rec >> 18) & 0x3FFFFu) - 3u, ob = (u32)(rec >> 36) & 0x3FFFFu;
        const u32 cl = kzg_enc_ll_code(T, ll), cm = kzg_enc_ml_code(T, mlb), co = (u32)kzg_highbit(ob);
        kzg_fsec_encode(w, sOF, T->st_of, T->dnb_of, T->dfs_of, co);
        kzg_fsec_encode(w, sML, T->st_ml, T->dnb_ml, T->dfs_ml, cm);
        kzg_fsec_encode(w, sLL, T->st_ll, T->dnb_ll, T->dfs_ll, cl);
        kzg_bw_add(w, ll, T->ll_tab[cl] >> 24);
        kzg_bw_flush(w);
        kzg_bw_add(w, mlb, T->ml_tab[cm] >> 24);
        kzg_bw_add(w, ob, co);
        kzg_bw_flush(w);
        if (w.ovf) return;                                       // not smaller than raw
    }
    kzg_bw_add(w, sML.value, 6);
    kzg_bw_add(w, sOF.value, 5);
    kzg_bw_add(w, sLL.value, 6);
    kzg_bw_add(w, 1, 1);                                         // end mark
    kzg_bw_flush(w);
    if (w.nbits) { if (w.p < w.end) *w.p = (u8)w.acc; else w.ovf = true; w.p++; }
    if (w.ovf || w.p >= limit) return;
    const u32 bsize = (u32)(w.p - blk);
    kzg_wr24(dst + hdr, 1u | (2u << 1) | (bsize << 3));
    ck.state = KZG_ENC_ST_COMPRESSED;
    ck.total = hdr + 3 + bsize;
}

#if defined(__CUDACC__)
namespace kzg_encd {

// unaligned 32-bit read from two aligned words
__device__ __forceinline__ u32 rd32u(const u8* p) {
    const u32* q = reinterpret_cast<const u32*>((uintptr_t)p & ~(uintptr_t)3);
    const u32 sh = ((u32)(uintptr_t)p & 3u) << 3;
    const u32 a = q[0];
    if (sh == 0) return a;
    return __funnelshift_r(a, q[1], sh);
}

// warp copy of n bytes, any alignment; reads stay inside [src & ~3, src + n + 3]
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

}  // namespace kzg_encd

// match stage: one warp per chunk, greedy parse with a single-probe hash table of 4-byte strings.
__global__ void __launch_bounds__(KZG_ENC_WARPS * 32) kzg_enc_match_kernel(KzgEncBatch B) {
    using namespace kzg_encd;
    extern __shared__ u32 enc_smem[];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32* table = enc_smem + (warp << KZG_ENC_HASH_LOG);
    for (;;) {
        u32 i = 0;
        if (lane == 0) i = atomicAdd(&B.tickets[0], 1u);
        i = __shfl_sync(0xFFFFFFFFu, i, 0);
        if (i >= B.n) break;
        const u32 n = (u32)B.src_size[i];
        KzgEncChunk ck;
        ck.n_seq = 0; ck.lit_size = 0; ck.state = 0; ck.total = 0;
        if (B.src_size[i] > KZG_BLOCK_MAX || B.dst_cap[i] < kzg_enc_bound(n) || n < 64) {     // errors / tiny chunks: raw block
            if (lane == 0) B.chunks[i] = ck;
            continue;
        }
        const u8* s = B.src + B.src_off[i];
        u8* lit = B.dst + B.dst_off[i] + kzg_enc_header_size(n) + 6;
        u64* seq = B.seq + (size_t)i * B.seq_cap;
        for (u32 k = lane; k < (1u << KZG_ENC_HASH_LOG); k += 32) table[k] = 0xFFFFFFFFu;
        __syncwarp();
        const u32 limit = n - 12;                // positions probed: p + 11 < n, so every read stays inside the chunk
        u32 pos = 0, anchor = 0, nseq = 0, nlit = 0;
        while (pos < limit) {
            const u32 p = pos + lane;
            const bool ok = p < limit;
            const u32 v = ok ? rd32u(s + p) : 0u;
            const u32 h = (v * 2654435761u) >> (32 - KZG_ENC_HASH_LOG);
            const u32 cand = ok ? table[h] : 0xFFFFFFFFu;
            const bool hit = cand < p && rd32u(s + cand) == v;
            const u32 m = __ballot_sync(0xFFFFFFFFu, hit);
            // positions up to the match start enter the table; the ones behind it are probed again by a later step
            // and must still find their older candidates then
            const u32 l = m ? __ffs(m) - 1 : 31u;
            if (ok && lane <= l) table[h] = p;
            __syncwarp();
            if (m == 0) { pos += 32; continue; }
            const u32 ms = pos + l, mc = __shfl_sync(0xFFFFFFFFu, cand, l);
            // forward extension, 128 bytes per step
            u32 len = 4;
            for (;;) {
                const u32 k = len + 4 * lane;
                const bool in = ms + k + 8 <= n;
                const u32 diff = in ? (rd32u(s + ms + k) ^ rd32u(s + mc + k)) : 1u;
                const u32 mm = __ballot_sync(0xFFFFFFFFu, diff != 0);
                if (mm) {
                    const u32 fl = __ffs(mm) - 1;
                    const u32 d = __shfl_sync(0xFFFFFFFFu, diff, fl);
                    const u32 inb = __shfl_sync(0xFFFFFFFFu, (u32)in, fl);
                    len += 4 * fl + (inb ? ((u32)__ffs(d) - 1) >> 3 : 0u);
                    break;
                }
                len += 128;
            }
            const u32 ll = ms - anchor;
            warp_copy(lit + nlit, s + anchor, ll, lane);
            if (lane == 0) seq[nseq] = (u64)ll | ((u64)len << 18) | ((u64)(ms - mc) << 36);
            nseq++;
            nlit += ll;
            anchor = pos = ms + len;
            if (nseq == B.seq_cap) break;
        }
        warp_copy_to_end(lit + nlit, s + anchor, n - anchor, lane);
        nlit += n - anchor;
        ck.n_seq = nseq; ck.lit_size = nlit;
        if (lane == 0) B.chunks[i] = ck;
    }
}

__global__ void __launch_bounds__(128) kzg_enc_encode_kernel(KzgEncBatch B) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B.n) kzg_enc_stage_encode(B, i);
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
