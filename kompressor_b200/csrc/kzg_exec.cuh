// Execute stage (s2 + s5) for sm_100a: raw / RLE blocks and ZSTD_execSequence, one warp per frame.
// Device-only code.
//
// A warp walks its frame's blocks in order and, inside a compressed block, the sequence records in
// batches of 32 (one lane per sequence).  Output is assembled in a per-warp shared-memory window and
// leaves for HBM only as 16-byte vector stores, 512 contiguous bytes per warp instruction:
//   * the window covers output positions [T0, T0 + KZG_EX_TILE) of the current block; when it is full its
//     completed 16-byte chunks are flushed and it slides, keeping the last KZG_EX_KEEP bytes as near history;
//   * every global read on the hot path is a cp.async (LDGSTS) into shared-memory staging, so no lane
//     ever waits on a load register: the literal bytes of a batch are contiguous in the literal stream
//     and are staged as 16-byte chunks one batch ahead; a match whose source has already left the window
//     (earlier in the block, earlier blocks of the frame: data this warp flushed to HBM/L2 before) stages
//     the 16-byte chunks covering its source, issued before the batch's literal pass and consumed after it;
//   * all per-lane copies are then shared -> shared, word-granular (aligned 32-bit loads + funnel shift),
//     bytes only at the ragged ends, with warp-uniform trip counts;
//   * matches whose source overlaps output produced by earlier lanes of the same batch wait for those
//     lanes: the batch is resolved in rounds, each round completing every lane whose source lies below
//     the first still-pending lane (usually one round);
//   * literal runs or matches longer than KZG_EX_LONG bytes are done by the whole warp, global -> global.
// No inter-warp synchronisation, no spinning.  What libzstd does here is ZSTD_execSequence, reached from
// the reference through ZSTD_decompressStream (kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:178).
#pragma once
#include "kzg_stages.cuh"

#define KZG_EX_WARPS 4
#ifndef KZG_EX_TILE
#define KZG_EX_TILE 3264u            // with the staging areas: 8 CTAs of 4 warps per SM (measured, CTAs x tile: 7 x 4096 -> 1.308 ms, 8 x 3264 -> 1.275, 9 x 2496 -> 1.82, 10 x 1856 -> 1.91: registers spill below 64)
#endif
#ifndef KZG_EX_KEEP
#define KZG_EX_KEEP 384u            // flushed history kept in the window after a slide (multiple of 16); measured with the 4 KiB window: 256 / 512 / 1024 / 1536 / 2048 -> execute 1.489 / 1.482 / 1.501 / 1.513 / 1.548 ms
#endif
#define KZG_EX_CAP (KZG_EX_TILE - KZG_EX_KEEP - 32u)   // most output one group of lanes may span
#define KZG_EX_LONG 64u             // literal runs / matches longer than this are copied by the whole warp
#ifndef KZG_EX_MEDLONG_HOT
#define KZG_EX_MEDLONG_HOT 0        // 1: the hot variant, too, keeps a medium-long match (<= 1 KiB, short literals) inside the window (measured: execute 1.420 -> 1.429 ms on logs, 1.231 -> 1.271 on JSON)
#endif
#ifndef KZG_EX_NO_LRING
#define KZG_EX_LRING 1              // literals are staged through a ring that runs ahead of the batches (else: batch by batch)
#endif
#ifdef KZG_EX_LRING
#define KZG_EX_LRING_BYTES 1024u    // the ring: four units
#define KZG_EX_LUNIT 256u           // one unit = 16 lanes x 16 bytes, staged when the batches have come within reach of it
#define KZG_EX_LMIRROR 96u          // the ring's first bytes once more behind its end: a run that starts before the end reads on there
#define KZG_EX_LAHEAD (512u + KZG_EX_LMIRROR)   // staged at the start of a batch: this much from its first literal (or to the end of the literals)
#define KZG_EX_LIT_BYTES (KZG_EX_LRING_BYTES + KZG_EX_LMIRROR)
#else
#define KZG_EX_LITSTAGE 1024u       // most literal bytes of one batch that are staged (else read directly)
#define KZG_EX_LIT_BYTES (KZG_EX_LITSTAGE + 32u)
#endif
#define KZG_EX_MSTAGE 80u           // per-lane match staging: the 16-byte chunks covering <= 64 source bytes
#define KZG_EX_WARP_SMEM (KZG_EX_TILE + KZG_EX_LIT_BYTES + 32u * KZG_EX_MSTAGE)
#define KZG_EX_SMEM (KZG_EX_WARPS * KZG_EX_WARP_SMEM + 32u)      // 16 bytes of slack on either side (over-reads of lane_copy_or)

namespace kzg_ex {

__device__ __forceinline__ u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ u32 lds8(u32 a) { u32 v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
// predicated forms: one instruction, no branch around it (a false predicate leaves `v` as passed in)
__device__ __forceinline__ u32 lds32_if(bool p, u32 a, u32 v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.shared.u32 %0, [%1];\n\t}" : "+r"(v) : "r"(a), "r"((u32)p) : "memory");
    return v;
}
__device__ __forceinline__ void sts32_if(bool p, u32 a, u32 v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}" :: "r"(a), "r"(v), "r"((u32)p) : "memory");
}
__device__ __forceinline__ void sts8(u32 a, u32 v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lds128(u32 a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(u32 a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async16(u32 dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
// Predicated form: one instruction, no branch around it.
__device__ __forceinline__ void cp_async16_if(bool p, u32 dst, const void* src) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" :: "r"(dst), "l"(src), "r"((u32)p) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
// History written earlier by this warp: read at L2 (the coherence point), never from a stale L1 line.
__device__ __forceinline__ u32 ldg32_hist(const u32* p) { return __ldcg(p); }
__device__ __forceinline__ u32 ldg8_hist(const u8* p) { return (u32)__ldcg(p); }

// One lane copies n (<= KZG_EX_LONG) bytes from shared address `sa` to shared address `d` (fill: the byte
// `fillv` instead).  Source and destination may be unaligned; a window -> window source must be at least
// 8 bytes behind its destination.  The source is read only as aligned words (combined by funnel shifts);
// the ragged head and tail of the destination are byte stores of an already assembled word, so no store
// waits on a byte load.  Up to two aligned words past the last source byte may be read (callers keep
// that inside their shared memory; the values are never used).
__device__ __forceinline__ void lane_copy(u32 d, u32 n, u32 sa, bool fill, u32 fillv) {
    if (n == 0) return;
    const u32 h = min(n, (0u - d) & 3u);
    u32 A = sa & ~3u, w0, w1;
    if (fill) w0 = w1 = fillv * 0x01010101u;
    else { w0 = lds32(A); w1 = lds32(A + 4); }
    {
        const u32 v = __funnelshift_r(w0, w1, (sa & 3u) << 3);
        if (h > 0) sts8(d, v);
        if (h > 1) sts8(d + 1, v >> 8);
        if (h > 2) sts8(d + 2, v >> 16);
    }
    d += h; n -= h;
    const u32 sb = (sa & 3u) + h;               // source byte offset inside w0:w1 after the head (0..6)
    const u32 r = (sb & 3u) << 3;
    u32 lo = sb >= 4u ? w1 : w0;
    A += sb >= 4u ? 4u : 0u;
    const u32 dend = d + (n & ~3u);
    if (fill) {
        for (; d != dend; d += 4) sts32(d, w0);
    } else {
        // up to four words per trip (the later ones predicated): fewer compares and branches per byte, and the loads
        // of a trip are independent of each other
        for (u32 rem = n >> 2; rem;) {
            const bool p2 = rem > 1, p3 = rem > 2, p4 = rem > 3;
            // (loads and stores alternate: a source 8 bytes behind its destination reads what this trip has written)
            const u32 h1 = lds32(A + 4);
            sts32(d, __funnelshift_r(lo, h1, r));
            const u32 h2 = lds32_if(p2, A + 8, 0u);
            sts32_if(p2, d + 4, __funnelshift_r(h1, h2, r));
            const u32 h3 = lds32_if(p3, A + 12, 0u);
            sts32_if(p3, d + 8, __funnelshift_r(h2, h3, r));
            const u32 h4 = lds32_if(p4, A + 16, 0u);
            sts32_if(p4, d + 12, __funnelshift_r(h3, h4, r));
            lo = p4 ? h4 : (p3 ? h3 : (p2 ? h2 : h1));
            const u32 done = min(rem, 4u);
            A += done << 2; d += done << 2; rem -= done;
        }
    }
    const u32 t = n & 3u;
    if (t) {
        const u32 hi = fill ? w0 : lds32(A + 4);
        const u32 v = __funnelshift_r(lo, hi, r);
        sts8(d, v);
        if (t > 1) sts8(d + 1, v >> 8);
        if (t > 2) sts8(d + 2, v >> 16);
    }
}

// Edge-free form of lane_copy for a window whose not yet written bytes are ZERO (kzg_exec_sequences keeps that
// invariant: the window is cleared at block start, behind every slide and after every warp-wide copy).  The copy
// walks the aligned destination words: word k takes the source bytes that land on it (two aligned loads, one funnel
// shift); full words are stored, and the first and the last word -- which may be shared with a neighbouring lane's
// output -- are OR-ed in with the bytes outside [d, d + n) masked off (RED.OR on shared memory).  No byte stores, no
// head / tail cases.  Reads up to two aligned words before the first and after the last source byte (callers keep
// that inside their shared memory; the values are masked off).  A window -> window source must be at least 8 bytes
// behind its destination: then every load of word k lies below destination word k, which this lane has written.
__device__ __forceinline__ void red_or32(u32 a, u32 v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void lane_copy_or(u32 d, u32 n, u32 sa, bool fill, u32 fillv) {
    if (n == 0) return;
    const u32 db = d & 3u, D = d - db;
    const u32 e = db + n;                                   // end of the copy, in bytes from D
    u32 rem = (e - 1u) >> 2;                                // destination words after the first
    const u32 m0 = 0xFFFFFFFFu << (db << 3);
    const u32 mL = 0xFFFFFFFFu >> (((0u - e) & 3u) << 3);
    if (fill) {
        const u32 v = fillv * 0x01010101u;
        red_or32(D, v & (rem ? m0 : (m0 & mL)));
        if (rem == 0) return;
        u32 dd = D + 4;
        for (--rem; rem; --rem, dd += 4) sts32(dd, v);
        red_or32(dd, v & mL);
        return;
    }
    const u32 s = sa - db;                                  // the source byte that lands on destination byte D
    u32 A = s & ~3u;
    const u32 r = (s & 3u) << 3;
    u32 lo = lds32(A), hi = lds32(A + 4);
    red_or32(D, __funnelshift_r(lo, hi, r) & (rem ? m0 : (m0 & mL)));
    if (rem == 0) return;
    lo = hi; A += 4;                                        // lo = the word at A
    u32 dd = D + 4;
    // full words, up to four per trip (the later ones predicated); loads and stores alternate: a source 8 bytes
    // behind its destination reads what this trip has written
    for (--rem; rem;) {
        const bool p2 = rem > 1, p3 = rem > 2, p4 = rem > 3;
        // (a load that is predicated off hands its predecessor's word on: the last word loaded ends up in h4)
        const u32 h1 = lds32(A + 4);
        sts32(dd, __funnelshift_r(lo, h1, r));
        const u32 h2 = lds32_if(p2, A + 8, h1);
        sts32_if(p2, dd + 4, __funnelshift_r(h1, h2, r));
        const u32 h3 = lds32_if(p3, A + 12, h2);
        sts32_if(p3, dd + 8, __funnelshift_r(h2, h3, r));
        const u32 h4 = lds32_if(p4, A + 16, h3);
        sts32_if(p4, dd + 12, __funnelshift_r(h3, h4, r));
        lo = h4;
        const u32 done = min(rem, 4u);
        A += done << 2; dd += done << 2; rem -= done;
    }
    // (measured and dropped: advancing the addresses by whole trips and loading the carried word again after the loop
    //  makes a trip 20 instructions instead of 24, and the kernel 17 % slower -- the reload sits on every lane's tail)
    hi = lds32(A + 4);
    red_or32(dd, __funnelshift_r(lo, hi, r) & mL);
}

// Straight-line form of lane_copy_or for n <= KZG_EX_LIT16 bytes (the literal run of nearly every sequence: 99 % are at
// most 16 bytes, while one lane in a batch of 32 usually has more than 8, which sent the whole warp through the loop
// above).  The copy touches at most five aligned destination words; all five are assembled (six aligned loads, five
// funnel shifts) and OR-ed in, the bytes outside [d, d + n) masked off.  No branches, no loop.  A lane with n == 0 must
// still pass addresses inside shared memory (loads and ORs are unconditional; it ORs zeros into five words from d on).
// Reads up to two aligned words before the first and 12 bytes after the last source byte, and ORs zeros into up to
// four words past the last destination byte (the callers' windows and staging areas allow for that).
#define KZG_EX_LIT16 16u
__device__ __forceinline__ void lane_copy_or16(u32 d, u32 n, u32 sa) {
    const u32 db = d & 3u, D = d - db;
    const u32 s = sa - db;                                  // the source byte that lands on destination byte D
    const u32 A = s & ~3u, r = (s & 3u) << 3;
    const i32 e8 = (i32)((db + n) << 3);                    // end of the copy, in bits from D (<= 152)
    const u32 w0 = lds32(A), w1 = lds32(A + 4), w2 = lds32(A + 8), w3 = lds32(A + 12), w4 = lds32(A + 16), w5 = lds32(A + 20);
    // low min(t, 32) bits set: the upper word of (0:~0) << t, shift clamped at 32.  Words past the end get the mask 0:
    // they are OR-ed all the same (a predicate would come back from ptxas as a branch around every word).
    const u32 m0 = (0xFFFFFFFFu << (db << 3)) & __funnelshift_lc(0xFFFFFFFFu, 0u, (u32)e8);
    red_or32(D, __funnelshift_r(w0, w1, r) & m0);
    red_or32(D + 4, __funnelshift_r(w1, w2, r) & __funnelshift_lc(0xFFFFFFFFu, 0u, (u32)max(e8 - 32, 0)));
    red_or32(D + 8, __funnelshift_r(w2, w3, r) & __funnelshift_lc(0xFFFFFFFFu, 0u, (u32)max(e8 - 64, 0)));
    red_or32(D + 12, __funnelshift_r(w3, w4, r) & __funnelshift_lc(0xFFFFFFFFu, 0u, (u32)max(e8 - 96, 0)));
    red_or32(D + 16, __funnelshift_r(w4, w5, r) & __funnelshift_lc(0xFFFFFFFFu, 0u, (u32)max(e8 - 128, 0)));
}

// Warp-cooperative global -> global copy of disjoint ranges, 16-byte vector stores.
// HIST: the source is output this warp wrote earlier (read at L2); otherwise read-only input data.
template <bool HIST>
__device__ __forceinline__ void warp_copy_global(u8* dst, const u8* src, u32 n, u32 lane) {
    const u32 head = min(n, (u32)((16 - ((uintptr_t)dst & 15)) & 15));
    for (u32 k = lane; k < head; k += 32) dst[k] = HIST ? (u8)ldg8_hist(src + k) : src[k];
    dst += head; src += head; n -= head;
    const u32 nvec = n >> 4;
    const u32 r = (u32)((uintptr_t)src & 3) * 8;
    const u32* sw = reinterpret_cast<const u32*>(src - ((uintptr_t)src & 3));
    for (u32 v = lane; v < nvec; v += 32) {
        u32 a, b, c, d, e = 0;
        if (HIST) { a = ldg32_hist(sw + 4 * v); b = ldg32_hist(sw + 4 * v + 1); c = ldg32_hist(sw + 4 * v + 2); d = ldg32_hist(sw + 4 * v + 3); if (r) e = ldg32_hist(sw + 4 * v + 4); }
        else { a = sw[4 * v]; b = sw[4 * v + 1]; c = sw[4 * v + 2]; d = sw[4 * v + 3]; if (r) e = sw[4 * v + 4]; }
        reinterpret_cast<uint4*>(dst)[v] = make_uint4(__funnelshift_r(a, b, r), __funnelshift_r(b, c, r),
                                                      __funnelshift_r(c, d, r), __funnelshift_r(d, e, r));
    }
    for (u32 k = (nvec << 4) + lane; k < n; k += 32) dst[k] = HIST ? (u8)ldg8_hist(src + k) : src[k];
}

__device__ __forceinline__ void warp_set_global(u8* dst, u32 byte, u32 n, u32 lane) {
    const u32 head = min(n, (u32)((16 - ((uintptr_t)dst & 15)) & 15));
    for (u32 k = lane; k < head; k += 32) dst[k] = (u8)byte;
    dst += head; n -= head;
    const u32 w = byte * 0x01010101u, nvec = n >> 4;
    for (u32 v = lane; v < nvec; v += 32) reinterpret_cast<uint4*>(dst)[v] = make_uint4(w, w, w, w);
    for (u32 k = (nvec << 4) + lane; k < n; k += 32) dst[k] = (u8)byte;
}

// The per-warp output window of the block being executed.
struct Win {
    u32 tile;           // shared address of window byte 0
    u8* out;            // global address of block position 0
    i32 T0;             // block position of window byte 0; (out + T0) is 16-byte aligned
    i32 Tv;             // window holds positions [Tv, cur)
    i32 F;              // positions below F are in global memory
    i64 boff;           // frame offset of block position 0
    const u8* dict_end; // one past the raw-content dictionary (history before the frame)
};

// History byte at block position p < w.Tv (already flushed): earlier output of the frame or the dictionary.
__device__ __forceinline__ u32 hist_byte(const Win& w, i32 p) {
    const i64 fi = w.boff + p;
    return fi >= 0 ? ldg8_hist(w.out + p) : (u32)w.dict_end[fi];
}

// Flushes positions [w.F, upto) as 16-byte vectors; with `all`, also the ragged tail (bytes).
__device__ __forceinline__ void flush(Win& w, i32 upto, bool all, u32 lane) {
    if (__builtin_expect(((uintptr_t)(w.out + w.F) & 15) != 0, 0)) {      // unaligned start (block start, or after a warp-wide copy)
        const i32 nh = min(upto - w.F, (i32)(16 - ((uintptr_t)(w.out + w.F) & 15)));
        if (!all && w.F + nh > upto) return;
        if ((i32)lane < nh) w.out[w.F + (i32)lane] = (u8)lds8(w.tile + (u32)(w.F - w.T0) + lane);
        w.F += nh;
    }
    const i32 nvec = (upto - w.F) >> 4;
    if (nvec > 0) {
        uint4* g = reinterpret_cast<uint4*>(w.out + w.F);
        const u32 t = w.tile + (u32)(w.F - w.T0);
        if ((i32)lane < nvec) g[lane] = lds128(t + (lane << 4));
        if (__builtin_expect(nvec > 32, 0)) {
#pragma unroll 1
            for (i32 v = (i32)lane + 32; v < nvec; v += 32) g[v] = lds128(t + ((u32)v << 4));
        }
        w.F += nvec << 4;
    }
    if (all && w.F < upto) {
        const i32 nt = upto - w.F;              // < 16
        if ((i32)lane < nt) w.out[w.F + (i32)lane] = (u8)lds8(w.tile + (u32)(w.F - w.T0) + lane);
        w.F = upto;
    }
}

// Clears window bytes [from, to) (from a multiple of 16): lane_copy_or ORs into unwritten bytes.
__device__ __forceinline__ void zero_window(u32 tile, u32 from, u32 to, u32 lane) {
    for (u32 x = from + (lane << 4); x < to; x += 512) sts128(tile + x, make_uint4(0, 0, 0, 0));
}

// Slides the window down, keeping KZG_EX_KEEP bytes of flushed history below w.F.
// Requires everything below `cur` (rounded down to 16) to be flushed already.
__device__ __forceinline__ void slide(Win& w, i32 cur, u32 lane) {
    i32 nT0 = w.F - (i32)KZG_EX_KEEP;
    if (nT0 <= w.T0) return;
    const u32 shift = (u32)(nT0 - w.T0);                    // multiple of 16
    const u32 nbytes = ((u32)(cur - nT0) + 15u) & ~15u;
    for (u32 o = 0; o < nbytes; o += 512) {
        const u32 x = o + (lane << 4);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (x < nbytes) v = lds128(w.tile + shift + x);
        __syncwarp();
        if (x < nbytes) sts128(w.tile + x, v);
        __syncwarp();
    }
    zero_window(w.tile, nbytes, nbytes + shift, lane);     // what the move vacated: unwritten bytes of the window are zero
    __syncwarp();
    w.T0 = nT0;
    if (w.Tv < nT0) w.Tv = nT0;
}

#ifdef KZG_EX_LRING
// Literal staging through a ring.  Stream coordinate x = (lit & 15) + l for literal l, so that x = 0 is a 16-byte
// aligned address; byte x lives at ring offset x & (KZG_EX_LRING_BYTES - 1).  Units of KZG_EX_LUNIT bytes are
// staged (asynchronously) from `lhi` up to `target`; chunks at or beyond x_end (the end of the literals) are not read.
// A unit at ring offset 0 is also copied behind the ring's end (its first KZG_EX_LMIRROR bytes).
__device__ __forceinline__ u32 stage_literal_units(u32 ring, const u8* lit_base16, u32 lhi, u32 target, u32 x_end, u32 lane) {
#pragma unroll 1
    for (; lhi < target; lhi += KZG_EX_LUNIT) {
        const u32 idx = lane & 15u;
        const u32 cx = lhi + (idx << 4);
        const bool mirror = lane >= 16u;
        const bool p = cx < x_end && (!mirror || ((lhi & (KZG_EX_LRING_BYTES - 1u)) == 0u && idx < KZG_EX_LMIRROR / 16u));
        const u32 dst = ring + (mirror ? KZG_EX_LRING_BYTES + (idx << 4) : (cx & (KZG_EX_LRING_BYTES - 1u)));
        cp_async16_if(p, dst, lit_base16 + cx);
    }
    return lhi;
}
#else
// Literal staging: the 16-byte chunks covering literal-stream bytes [l0, l1) are copied (asynchronously)
// to `stage`; byte l of the stream then sits at stage + (l - l0) + ((lit + l0) & 15).
// Returns false (nothing issued) if the range is too large to stage.
__device__ __forceinline__ bool stage_literals(u32 stage, const u8* lit, u32 l0, u32 l1, u32 lane) {
    if (l1 - l0 > KZG_EX_LITSTAGE) return false;
    const u8* g0 = lit + l0;
    const u32 a = (u32)((uintptr_t)g0 & 15);
    const u32 nch = (a + (l1 - l0) + 15u) >> 4;             // <= 66
    g0 -= a;
    cp_async16_if(lane < nch, stage + (lane << 4), g0 + (lane << 4));
    if (__builtin_expect(nch > 32, 0)) {
#pragma unroll 1
        for (u32 c = lane + 32; c < nch; c += 32) cp_async16(stage + (c << 4), g0 + ((size_t)c << 4));
    }
    return true;
}
#endif

}  // namespace kzg_ex

// Executes one compressed block with sequences. Returns (uniformly) 0 = done, 1 = the block is corrupt, 2 = the
// block needs the FULL variant.  smem: this warp's shared memory (window | literal staging | match staging).
template <bool FULL>
__device__ __forceinline__ u32 kzg_exec_sequences(const KzgBatch& B, const KzgBlock& blk, const KzgCBlock* cb, u8* out,
                                                   u32 smem, u32 lane) {
    using namespace kzg_ex;
    const u32 nseq = cb->n_seq, lit_regen = cb->lit_regen, lit_type = cb->lit_type;
    const u8* lit = lit_type >= 2 ? B.lit_pool + cb->lit_out : B.src + cb->src + cb->lit_src;
    const bool lfill = lit_type == 1;
    const bool bulk_lits = cb->bulk_lits != 0;      // long literal runs of this block were copied by kzg_bulk_kernel
    const u32 rle = lfill ? (u32)lit[0] : 0u;
    const u32 lit_align = (u32)((uintptr_t)lit & 15);
    const u64* seqs = B.seq_pool + cb->seq_out;
    const u32 ri0 = cb->rep_init[0], ri1 = cb->rep_init[1], ri2 = cb->rep_init[2];
    // history available before block position 0 (earlier blocks + dictionary), clamped for 32-bit compares
    const u64 hist64 = blk.out_off + B.dict_size;
    const u32 hist_before_block = hist64 > 0x40000000ull ? 0x40000000u : (u32)hist64;
    const u32 lstage = smem + KZG_EX_TILE;
    const u32 mstage = smem + KZG_EX_TILE + KZG_EX_LIT_BYTES + lane * KZG_EX_MSTAGE;
    Win w;
    w.tile = smem; w.out = out;
    w.T0 = -(i32)((uintptr_t)out & 15);
    w.Tv = 0; w.F = 0;
    w.boff = (i64)blk.out_off;
    w.dict_end = B.dict + B.dict_size;
    zero_window(w.tile, 0u, KZG_EX_TILE, lane);
    __syncwarp();
    bool err = false, defer = false;
    u64 carry = 0;                       // last record of the previous batch
    u64 rec_next = lane < nseq ? seqs[lane] : 0ull;
    const u64* rec_ptr = seqs + lane;
    i32 cur = 0;
#ifdef KZG_EX_LRING
    // literal staging: the ring starts KZG_EX_LAHEAD bytes (whole units) ahead of the first batch
    const u8* const lit_base16 = lit - lit_align;
    const u32 lit_x_end = lfill ? 0u : lit_align + lit_regen;
    u32 lhi = stage_literal_units(lstage, lit_base16, 0u, (lit_align + KZG_EX_LAHEAD + KZG_EX_LUNIT - 1u) & ~(KZG_EX_LUNIT - 1u), lit_x_end, lane);
#else
    // literal staging of the first batch
    bool lit_staged = false;
    if (!lfill) {
        const u32 nv = min(32u, nseq);
        const u32 l1 = (u32)(__shfl_sync(0xFFFFFFFFu, rec_next, nv - 1) >> 28) & 0x3FFFFu;
        lit_staged = stage_literals(lstage, lit, 0u, l1, lane);
    }
#endif
    cp_commit();

    for (u32 b0 = 0; b0 < nseq; b0 += 32) {
        const u32 i = b0 + lane;
        const bool valid = i < nseq;
        const u64 rec = rec_next;
        rec_ptr += 32;
        rec_next = (i + 32 < nseq) ? *rec_ptr : 0ull;              // prefetch the next batch
        u64 prev = __shfl_up_sync(0xFFFFFFFFu, rec, 1);
        if (lane == 0) prev = carry;
        const u32 nvalid = min(32u, nseq - b0);
        const u32 this_l0 = (u32)(carry >> 28) & 0x3FFFFu;        // literal position the staging area starts at
        carry = __shfl_sync(0xFFFFFFFFu, rec, nvalid - 1);
        const u32 lit_end = (u32)(rec >> 28) & 0x3FFFFu, p_lit_end = (u32)(prev >> 28) & 0x3FFFFu;
        const u32 out_end = (u32)(rec >> 46), p_out_end = (u32)(prev >> 46);
        const u32 ll = valid ? lit_end - p_lit_end : 0u;
        const u32 mdst = p_out_end + ll;
        const u32 ml = valid ? out_end - mdst : 0u;
        u32 off = (u32)rec & KZG_OFF_MASK;
        if (off & KZG_OFF_SYM) {
            const u32 k = off & 3, dd = (off & (KZG_OFF_SYM - 1)) >> 2;
            const u32 base = k == 0 ? ri0 : (k == 1 ? ri1 : ri2);
            off = (base >= KZG_OFF_INVALID || base <= dd) ? KZG_OFF_INVALID : base - dd;
        }
        // (the sequence kernel guarantees non-decreasing lit_end / out_end within the block limits)
        const bool bad = valid && (off - 1u >= KZG_OFF_INVALID - 1u || off > hist_before_block + mdst);
        const bool islong = ll > KZG_EX_LONG || ml > KZG_EX_LONG;
        const u32 longm = __ballot_sync(0xFFFFFFFFu, islong);
        if (__builtin_expect(__any_sync(0xFFFFFFFFu, bad), 0)) { err = true; break; }
        const i32 spos = (i32)mdst - (i32)off;
        const i32 be = (i32)((u32)(carry >> 46));
        // make room for the whole batch first if it fits a window (the usual case), so that the source
        // classification below sees the window it will run with
        if (__builtin_expect(longm == 0 && (u32)(be - w.T0) > KZG_EX_TILE, 0)) { flush(w, cur, false, lane); slide(w, cur, lane); }

        // ---- matches whose source is not (entirely) inside the window any more but has been flushed: stage the
        // 16-byte chunks covering it now.  (Tv and F only grow, so the classification stays true.)
        bool gsrc = !islong && ml > 0 && spos < w.Tv && spos + (i32)ml <= w.F && w.boff + spos >= 0;
        const u32 gsa = mstage + (u32)((uintptr_t)(out + spos) & 15);
        auto stage_sources = [&](bool want) {
            const u8* g = out + spos;
            const u32 a = (u32)((uintptr_t)g & 15);
            const u32 nch = want ? (a + ml + 15u) >> 4 : 0u;             // <= 5
            g -= a;
            cp_async16_if(nch > 0, mstage, g);
            cp_async16_if(nch > 1, mstage + 16, g + 16);
            cp_async16_if(nch > 2, mstage + 32, g + 32);
            cp_async16_if(nch > 3, mstage + 48, g + 48);
            cp_async16_if(nch > 4, mstage + 64, g + 64);
        };
        stage_sources(gsrc);
        cp_commit();
        // literal bytes of this batch were staged one batch ago: all groups but the newest are complete
        cp_wait<1>();
        __syncwarp();
#ifdef KZG_EX_LRING
        // every literal of the batch is staged if its last one is (staged: KZG_EX_LAHEAD from the first one, or all there are)
        const bool this_staged = !lfill && lit_align + ((u32)(carry >> 28) & 0x3FFFFu) <= lhi;
        const u32 lsrc = lstage + ((lit_align + p_lit_end) & (KZG_EX_LRING_BYTES - 1u));   // this lane's literals
        (void)this_l0;
#else
        const bool this_staged = lit_staged;
        const u32 lsrc = lstage + ((lit_align + this_l0) & 15u) - this_l0 + p_lit_end;   // this lane's literals
#endif
        bool first_round = true;

        // One group of lanes (`mine`), whose output ends at block position ge and fits the window.
        auto run_group = [&](bool mine, i32 ge) {
            // ---- literal runs of the group (no dependencies)
            const u32 gll = mine ? ll : 0u;
#ifndef KZG_EX_NO_LIT16
            if (__builtin_expect(this_staged && !__any_sync(0xFFFFFFFFu, gll > KZG_EX_LIT16), 1)) {
                // (lanes without literals OR zeros into a word of their own at the start of the window and read the staging area)
                lane_copy_or16(gll ? w.tile + (u32)((i32)p_out_end - w.T0) : w.tile + (lane << 2), gll, gll ? lsrc : lstage);
            } else
#endif
            if (__builtin_expect(this_staged || lfill, 1)) {
                lane_copy_or(w.tile + (u32)((i32)p_out_end - w.T0), gll, lsrc, lfill, rle);
            } else {                 // literals beyond what is staged (rare): those runs come straight from global memory
#ifdef KZG_EX_LRING
                const bool st = lit_align + lit_end <= lhi;
                lane_copy_or(w.tile + (u32)((i32)p_out_end - w.T0), st ? gll : 0u, lsrc, false, 0u);
                __syncwarp();
#else
                const bool st = false;
#endif
                if (mine && !st) {
                    const u32 d = w.tile + (u32)((i32)p_out_end - w.T0);
                    for (u32 k = 0; k < ll; k++) sts8(d + k, (u32)__ldg(lit + p_lit_end + k));
                }
            }
            __syncwarp();
            // ---- matches, in dependency rounds
            bool pending = mine && ml > 0;
            // source bytes this match needs to be final: everything below `need`
            const i32 need = off >= ml ? spos + (i32)ml : (i32)mdst;
            const bool slow = !gsrc && !(spos >= w.Tv && off >= 8);
            const bool anyslow = __any_sync(0xFFFFFFFFu, pending && slow);
            const u32 sa = gsrc ? gsa : w.tile + (u32)(spos - w.T0);
            const u32 da = w.tile + (u32)((i32)mdst - w.T0);
            for (;;) {
                const u32 pm = __ballot_sync(0xFFFFFFFFu, pending);
                if (!pm) break;
                const i32 fin = (i32)__shfl_sync(0xFFFFFFFFu, mdst, __ffs(pm) - 1);   // all below is final
                const bool go = pending && need <= fin;
                // (measured and dropped: letting the whole warp copy the one or two matches a later round usually completes,
                //  a destination word per lane -- ~20 instructions against ~75 for the per-lane pass -- made the kernel 3 % slower)
                if (first_round) { cp_wait<0>(); first_round = false; }     // staged match sources have landed
                lane_copy_or(da, (go && !slow) ? ml : 0u, sa, false, 0u);
                if (__builtin_expect(anyslow, 0)) {
                    __syncwarp();
                    if (go && slow) {     // short period, or a source straddling the flushed / window boundary: bytes in order
                        for (u32 k = 0; k < ml; k++) {
                            const i32 p = spos + (i32)k;
                            const u32 v = p >= w.Tv ? lds8(w.tile + (u32)(p - w.T0)) : hist_byte(w, p);
                            sts8(da + k, v);
                        }
                    }
                }
                pending = pending && !go;
                __syncwarp();
            }
            cur = ge;
            // (the window is flushed when it has to slide, ~3.5 KiB at a time; flushing every fourth or eighth batch as well
            //  was measured: execute 1.482 / 1.490 ms against 1.472 without)
        };

        // After the window has slid or started over in the middle of a batch (below), sources of lanes still to run may
        // have left it: they are staged from global memory now, like the ones that had left it at the start of the batch
        // (else those lanes copy byte by byte from L2, one dependent load per byte: measured at 5 % of the kernel's stall
        // samples on log text, where one batch in fifty holds a long match).
        auto restage = [&](u32 left) {
            const bool g2 = ((left >> lane) & 1u) && !gsrc && !islong && ml > 0 && spos < w.Tv && spos + (i32)ml <= w.F && w.boff + spos >= 0;
            if (__any_sync(0xFFFFFFFFu, g2)) {
                stage_sources(g2);
                cp_commit();
                cp_wait<0>();
                __syncwarp();
                gsrc = gsrc || g2;
            }
        };

        if (__builtin_expect(longm == 0 && (u32)(be - cur) <= KZG_EX_CAP, 1)) {
            run_group(valid, be);                       // the usual case: the whole batch at once
        } else {
            u32 todo = nvalid == 32 ? 0xFFFFFFFFu : ((1u << nvalid) - 1u);     // lanes not yet executed, in order
            while (todo) {
                // next group: lanes up to the next long sequence whose output still fits the window
                const u32 firstl = __ffs(todo) - 1;
                const i32 gs = (i32)__shfl_sync(0xFFFFFFFFu, p_out_end, firstl);
                const u32 longs = longm & todo;
                const u32 stop = longs ? (u32)__ffs(longs) - 1 : 32u;      // lane index of the next long sequence
                const u32 fitm = __ballot_sync(0xFFFFFFFFu, (u32)((i32)out_end - gs) <= KZG_EX_CAP);
                u32 grp = stop >= 32 ? todo : (todo & ((1u << stop) - 1u));
                {   // keep the leading run of lanes that fit
                    const u32 nofit = grp & ~fitm;
                    if (nofit) grp &= (1u << (__ffs(nofit) - 1)) - 1u;
                }
                if (grp) {
                    const u32 lastl = 31 - __clz(grp);
                    const i32 ge = (i32)__shfl_sync(0xFFFFFFFFu, out_end, lastl);
                    if ((u32)(ge - w.T0) > KZG_EX_TILE) { flush(w, cur, false, lane); slide(w, cur, lane); restage(todo); }
                    run_group((grp >> lane) & 1u, ge);
                    todo &= ~grp;
                } else {
                    // the next sequence in order is a long one: the warp copies it straight to global memory
                    const u32 l_ll = __shfl_sync(0xFFFFFFFFu, ll, stop), l_ml = __shfl_sync(0xFFFFFFFFu, ml, stop);
                    const u32 l_ple = __shfl_sync(0xFFFFFFFFu, p_lit_end, stop), l_poe = __shfl_sync(0xFFFFFFFFu, p_out_end, stop);
                    const u32 l_off = __shfl_sync(0xFFFFFFFFu, off, stop);
                    // (FULL variant) A medium-long match with short literals stays inside the window: the lane copies its
                    // literals, the warp copies the match 32 bytes per step (a source at least 32 bytes behind, or disjoint).
                    {
                        const i32 sp0 = (i32)(l_poe + l_ll) - (i32)l_off;
                        if ((FULL || KZG_EX_MEDLONG_HOT) && l_ll <= KZG_EX_LONG && (this_staged || lfill) && l_ml <= 1024u && (l_off >= 32u || l_off >= l_ml) && w.boff + sp0 >= 0) {
                            const i32 end = (i32)(l_poe + l_ll + l_ml);
                            if ((u32)(end - w.T0) > KZG_EX_TILE) { flush(w, cur, false, lane); slide(w, cur, lane); restage(todo & ~(1u << stop)); }
                            const bool in_win = sp0 >= w.Tv, in_glob = sp0 + (i32)l_ml <= w.F;
                            if ((u32)(end - w.T0) <= KZG_EX_TILE && (in_win || in_glob)) {
                                lane_copy_or(w.tile + (u32)((i32)l_poe - w.T0), lane == stop ? ll : 0u, lsrc, lfill, rle);
                                __syncwarp();
                                const u32 da = w.tile + (u32)((i32)(l_poe + l_ll) - w.T0);
                                if (in_win) {
                                    const u32 sa = w.tile + (u32)(sp0 - w.T0);
                                    for (u32 k0 = 0; k0 < l_ml; k0 += 32) {
                                        const u32 k = k0 + lane;
                                        const u32 v = k < l_ml ? lds8(sa + k) : 0u;
                                        if (k < l_ml) sts8(da + k, v);
                                        __syncwarp();
                                    }
                                } else {
                                    for (u32 k = lane; k < l_ml; k += 32) sts8(da + k, ldg8_hist(out + sp0 + (i32)k));
                                    __syncwarp();
                                }
                                cur = end;
                                todo &= ~(1u << stop);
                                continue;
                            }
                        }
                    }
                    if (!FULL && l_off < l_ml && l_ml >= 1024u) { defer = true; break; }      // a long run: left to the FULL variant
                    flush(w, cur, true, lane);
                    __syncwarp();
                    u8* d = out + l_poe;
                    if (bulk_lits && l_ll >= KZG_BULK_MIN) { /* written by kzg_bulk_kernel */ }
                    else if (lfill) warp_set_global(d, rle, l_ll, lane);
                    else warp_copy_global<false>(d, lit + l_ple, l_ll, lane);
                    __threadfence_block();
                    __syncwarp();
                    d += l_ll;
                    const i32 sp = (i32)(l_poe + l_ll) - (i32)l_off;
                    if (l_off >= l_ml && w.boff + sp >= 0) {
                        warp_copy_global<true>(d, out + sp, l_ml, lane);      // disjoint source inside the frame
                    } else if (FULL && l_ml >= 1024u && w.boff + sp >= 0 && l_off >= 64) {
                        // overlapping with a long period: pieces of l_off bytes, each disjoint from its (already
                        // written) source
                        for (u32 done = 0; done < l_ml; done += l_off) {
                            warp_copy_global<true>(d + done, out + sp + done, min(l_off, l_ml - done), lane);
                            __threadfence_block();
                            __syncwarp();
                        }
                    } else if (FULL && l_ml >= 1024u && w.boff + sp >= 0) {
                        // overlapping with a short period (runs): expand the period to ~1 KiB in the (flushed, hence free)
                        // window, then stream it out as 16-byte vectors; positions wrap by compare-subtract, no division
                        const u32 L = l_off * (1024u / l_off);
                        for (u32 k = lane; k < L; k += 32) sts8(w.tile + k, ldg8_hist(out + sp + (i32)(k % l_off)));
                        __syncwarp();
                        const u32 head = min(l_ml, (u32)((16 - ((uintptr_t)d & 15)) & 15));
                        if (lane < head) d[lane] = (u8)lds8(w.tile + lane);                 // head < 16 <= L
                        const u32 nvec = (l_ml - head) >> 4;
                        u32 p = (head + (lane << 4)) % L;                                   // pattern phase of this lane's chunk
                        const u32 stepw = 512u % L;                                         // phase advance per pass of 32 chunks
                        for (u32 v = lane; v < nvec; v += 32) {
                            u32 wv[4];
                            u32 q = p;
#pragma unroll
                            for (int wi = 0; wi < 4; wi++) {
                                u32 x = 0;
#pragma unroll
                                for (int bi = 0; bi < 4; bi++) { x |= lds8(w.tile + q) << (8 * bi); q++; q = q >= L ? q - L : q; }
                                wv[wi] = x;
                            }
                            reinterpret_cast<uint4*>(d + head)[v] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                            p += stepw; p = p >= L ? p - L : p;
                        }
                        const u32 donev = head + (nvec << 4);
                        if (lane < l_ml - donev) d[donev + lane] = (u8)lds8(w.tile + (donev + lane) % L);
                    } else {
                        // a short overlapping match, or one reaching into the dictionary: every byte equals one of the
                        // l_off bytes just before the match, all of which are final
                        for (u32 k = lane; k < l_ml; k += 32) {
                            const i32 p = sp + (i32)(l_off >= l_ml ? k : k % l_off);
                            const i64 fi = w.boff + p;
                            d[k] = fi >= 0 ? (u8)ldg8_hist(out + p) : w.dict_end[fi];
                        }
                    }
                    __threadfence_block();
                    __syncwarp();
                    cur = (i32)(l_poe + l_ll + l_ml);
                    w.F = w.Tv = cur;
                    w.T0 = cur - (i32)((uintptr_t)(out + cur) & 15);
                    zero_window(w.tile, 0u, KZG_EX_TILE, lane);     // the window starts over, empty
                    __syncwarp();
                    todo &= ~(1u << stop);
                    restage(todo);
                }
            }
        }
        if (defer) break;
        // ---- stage the next batch's literals (this batch no longer reads the staging area)
        cp_wait<0>();
        __syncwarp();
#ifdef KZG_EX_LRING
        {   // keep the ring KZG_EX_LAHEAD ahead of the next batch's first literal, as far as the bytes still needed allow
            // (the units below that literal's are free); usually nothing to do, a unit every fourth batch
            const u32 x0 = lit_align + ((u32)(carry >> 28) & 0x3FFFFu);
            const u32 target = min((x0 + KZG_EX_LAHEAD + KZG_EX_LUNIT - 1u) & ~(KZG_EX_LUNIT - 1u), (x0 & ~(KZG_EX_LUNIT - 1u)) + KZG_EX_LRING_BYTES);
            if (__builtin_expect(lhi < target && lhi < lit_x_end && b0 + 32 < nseq, 0)) {
                // (after a long literal run the units in between are dead already: not staged)
                lhi = stage_literal_units(lstage, lit_base16, max(lhi, x0 & ~(KZG_EX_LUNIT - 1u)), target, lit_x_end, lane);
            }
        }
#else
        lit_staged = false;
        if (!lfill && b0 + 32 < nseq) {
            const u32 nv = min(32u, nseq - b0 - 32);
            const u32 l0 = (u32)(carry >> 28) & 0x3FFFFu;
            const u32 l1 = (u32)(__shfl_sync(0xFFFFFFFFu, rec_next, nv - 1) >> 28) & 0x3FFFFu;
            if (l1 >= l0) lit_staged = stage_literals(lstage, lit, l0, l1, lane);
        }
#endif
        cp_commit();
    }
    cp_wait<0>();
    __syncwarp();
    if (err) return 1u;
    if (defer) return 2u;
    // tail: what is left in the window, then the literals after the last sequence
    flush(w, cur, true, lane);
    __syncwarp();
    {
        const u32 used = (u32)(carry >> 28) & 0x3FFFFu;
        if (lit_regen > used && lit_regen - used < KZG_BULK_MIN) {          // longer tails: kzg_bulk_kernel
            if (lfill) warp_set_global(out + cur, rle, lit_regen - used, lane);
            else warp_copy_global<false>(out + cur, lit + used, lit_regen - used, lane);
        }
    }
    __threadfence_block();
    __syncwarp();
    return 0u;
}

// FULL = false: the hot variant.  It has no code for long overlapping matches (runs); a frame that contains one is
// put on the deferred list and executed again, from its first block, by the FULL = true variant, which is launched
// right after over that list only.  Keeping the rare path out of the hot kernel is worth ~5 % of its speed.
#define KZG_EX_BIG_FRAME_BLOCKS 8u
template <bool FULL>
#ifndef KZG_EX_MINBLOCKS
#define KZG_EX_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(KZG_EX_WARPS * 32, KZG_EX_MINBLOCKS) kzg_exec_kernel(KzgBatch B) {
    using namespace kzg_ex;
    extern __shared__ uint4 smem4[];
    u32 lane = threadIdx.x & 31;
    asm volatile("" : "+r"(lane));       // (in a register, like the base below: ptxas otherwise re-reads SR_TID.X eight times per batch; execute 1.270 -> 1.255 ms)
    const u32 warp = threadIdx.x >> 5;
    u32 smem = (u32)__cvta_generic_to_shared(smem4) + 16u + warp * KZG_EX_WARP_SMEM;
    asm volatile("" : "+r"(smem));       // keep the base in a register: ptxas otherwise rebuilds it (S2R, LEA, IMAD) twice per batch (execute 1.53 -> 1.50 ms)
    if (B.ctr->overflow & 3u) return;
    for (;;) {
        u32 f = 0;
        if (lane == 0) f = atomicAdd(FULL ? &B.ctr->ticket_exec2 : &B.ctr->ticket_exec, 1u);
        f = __shfl_sync(0xFFFFFFFFu, f, 0);
        if (f >= (FULL ? B.ctr->n_defer : B.n_frames)) break;
        if (FULL) {
            f = B.dlist[f];
            if (f & KZG_DEFER_BIG) continue;            // a frame of many blocks: kzg_exec4_kernel's
        }
        const KzgFrame* fr = &B.frames[f];
        if (fr->status != KZG_OK) continue;
        u8* const fdst = B.dst + B.dst_off[f];
        const u32 nblk = fr->n_blocks, first = fr->first_block;
        bool failed = false;
        // frames of many blocks (large windows, long matches) go straight to the full variant, which keeps medium-long
        // matches inside the window
        bool deferred = !FULL && nblk > KZG_EX_BIG_FRAME_BLOCKS;
        u32 kfail = 0;
        for (u32 k = 0; k < nblk && !failed && !deferred; k++) {
            const KzgBlock blk = B.blocks[first + k];
            u8* out = fdst + blk.out_off;
            // raw / RLE blocks and literal-only blocks were written by kzg_bulk_kernel
            if (blk.type == 2) {
                const KzgCBlock* cb = &B.cblocks[blk.cidx];
                if (cb->n_seq) {
                    const u32 r = kzg_exec_sequences<FULL>(B, blk, cb, out, smem, lane);
                    failed = r == 1;
                    deferred = r == 2;
                    kfail = k;
                }
            }
            __threadfence_block();
            __syncwarp();
        }
        if (failed && lane == 0) { B.frames[f].status = KZG_E_CORRUPTION; B.frames[f].fail_block = kfail; }
        // the full variant redoes a frame with long runs; a frame of many blocks goes to the CTA-per-frame kernel
        if (deferred && lane == 0) B.dlist[atomicAdd(&B.ctr->n_defer, 1u)] = f | ((B.big_by_cta && nblk > KZG_EX_BIG_FRAME_BLOCKS) ? KZG_DEFER_BIG : 0u);
    }
}
