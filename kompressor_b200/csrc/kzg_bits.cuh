// Bit readers. One GPU thread owns one reader; the same code compiles for the host (tests/hostsim).
#pragma once
#include "kzg_common.cuh"
#if !defined(__CUDA_ARCH__)
#include <string.h>
#endif

// 16-byte aligned chunk load with the bits that precede the stream zeroed (Zstandard backward
// bitstreams read as zero past their first byte).  Requires the 16-byte
// aligned chunk holding any valid byte to be readable (cudaMalloc guarantees this).
KZG_HD void kzg_ld_chunk(const u8* base16, i32 ci, u32 lowbits, u32& a, u32& b, u32& c, u32& d) {
    if (ci < 0) { a = b = c = d = 0; return; }
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4*>(base16 + ((size_t)ci << 4));
    a = v.x; b = v.y; c = v.z; d = v.w;
#else
    u32 t[4];
    memcpy(t, base16 + ((size_t)ci << 4), 16);
    a = t[0]; b = t[1]; c = t[2]; d = t[3];
#endif
    if (ci == 0 && lowbits) {
        if (lowbits >= 32) a = 0; else a &= ~0u << lowbits;
        if (lowbits >= 64) b = 0; else if (lowbits > 32) b &= ~0u << (lowbits - 32);
        if (lowbits >= 96) c = 0; else if (lowbits > 64) c &= ~0u << (lowbits - 64);
        if (lowbits > 96) d &= ~0u << (lowbits - 96);
    }
}

KZG_HD u32 kzg_ld_word(const u8* base16, i32 wi, u32 lowbits) {
    if (wi < 0) return 0;
    u32 v;
#if defined(__CUDA_ARCH__)
    v = *reinterpret_cast<const u32*>(base16 + ((size_t)wi << 2));
#else
    memcpy(&v, base16 + ((size_t)wi << 2), 4);
#endif
    const u32 lo = (u32)wi << 5;
    if (lo < lowbits) v = (lowbits - lo >= 32) ? 0u : (v & (~0u << (lowbits - lo)));
    return v;
}

// Backward bitstream reader (Huffman streams, FSE-coded Huffman weights, sequences).
// `win` holds the next `wbits` unread bits left-aligned; refill() appends exactly 32 bits, so the
// funnel-shift amount `sh` is a per-stream constant and every refill costs one register pop.
// Source words are fetched as aligned 16-byte chunks, one chunk ahead of use.
struct KzgBwd {
    const u8* base;
    u64 win;
    i64 left;          // unread bits of the stream; negative once a read ran past the first byte
    u32 w0, w1, w2, w3;
    u32 p0, p1, p2, p3;
    u32 hi, sh, lowbits;
    i32 wbits, nw, ci;

    // false if the stream is empty or its last byte is zero (no padding marker) => corruption
    KZG_HD bool init(const u8* start, u32 size) {
        win = 0; wbits = 0; left = 0;
        if (size == 0) return false;
        const u32 last = start[size - 1];
        if (last == 0) return false;
        const u32 T = 8u * (size - 1) + (u32)kzg_highbit(last);
        base = (const u8*)((uintptr_t)start & ~(uintptr_t)15);
        lowbits = 8u * (u32)(start - base);
        const u32 Tp = T + lowbits;
        sh = Tp & 31u;
        const i32 nwi = (i32)(Tp >> 5) - 1;
        hi = sh ? kzg_ld_word(base, nwi + 1, lowbits) : 0u;
        const i32 cq = nwi >> 2;
        kzg_ld_chunk(base, cq, lowbits, w0, w1, w2, w3);
        nw = (nwi & 3) + 1;
        for (i32 k = nw; k < 4; k++) { w3 = w2; w2 = w1; w1 = w0; w0 = 0; }
        ci = cq - 1;
        kzg_ld_chunk(base, ci, lowbits, p0, p1, p2, p3);
        left = (i64)T;
        return true;
    }
    KZG_HD void refill() {     // requires wbits <= 32
        const u32 lo = w3;
        w3 = w2; w2 = w1; w1 = w0;
        if (--nw == 0) {
            w3 = p3; w2 = p2; w1 = p1; w0 = p0; nw = 4;
            ci--;
            kzg_ld_chunk(base, ci, lowbits, p0, p1, p2, p3);
        }
        const u32 piece = kzg_funnel_r(lo, hi, sh);
        hi = lo;
        win |= (u64)piece << (32 - wbits);
        wbits += 32;
    }
    KZG_HD void ensure32() { if (wbits <= 32) refill(); }            // afterwards wbits > 32
    KZG_HD u32 peek(u32 n) const { return (u32)((win >> 1) >> (63 - n)); }   // 0 <= n <= 32 <= wbits
    KZG_HD void skip(u32 n) { win <<= n; wbits -= (i32)n; left -= (i64)n; }  // n < 64
    KZG_HD u32 read(u32 n) { const u32 v = peek(n); skip(n); return v; }
};

// Backward reader for the long-running per-thread decoders (sequence state machine, Huffman streams).
// Same bit semantics as KzgBwd, but global loads are issued ONLY in advance(), which every lane of a
// warp calls at the same program point (every R symbols).  The register scoreboard is per warp: a
// load issued by one lane into a register blocks every other lane's next read of that register, so
// per-lane "load when my queue runs dry" prefetching stalls the whole warp for a full memory latency
// on almost every iteration.  Here chunks travel  x,y (load targets) -> a,b (private) -> c (current);
// x,y are only read in advance(), i.e. at least R symbols after the newest load into them.
struct KzgBwdQ {
    const u8* base;
    u64 win;
    i64 left;
    u32 c0, c1, c2, c3;
    u32 a0, a1, a2, a3, b0, b1, b2, b3;
    u32 x0, x1, x2, x3, y0, y1, y2, y3;
    u32 hi, sh, lowbits, valid;          // valid: bit0 a, bit1 b, bit2 x, bit3 y
    i32 wbits, nw, ci;

    KZG_HD bool init(const u8* start, u32 size) {
        win = 0; wbits = 0; left = 0; valid = 0;
        if (size == 0) return false;
        const u32 last = start[size - 1];
        if (last == 0) return false;
        const u32 T = 8u * (size - 1) + (u32)kzg_highbit(last);
        base = (const u8*)((uintptr_t)start & ~(uintptr_t)15);
        lowbits = 8u * (u32)(start - base);
        const u32 Tp = T + lowbits;
        sh = Tp & 31u;
        const i32 nwi = (i32)(Tp >> 5) - 1;
        hi = sh ? kzg_ld_word(base, nwi + 1, lowbits) : 0u;
        const i32 cq = nwi >> 2;
        kzg_ld_chunk(base, cq, lowbits, c0, c1, c2, c3);
        kzg_ld_chunk(base, cq - 1, lowbits, a0, a1, a2, a3);
        kzg_ld_chunk(base, cq - 2, lowbits, b0, b1, b2, b3);
        kzg_ld_chunk(base, cq - 3, lowbits, x0, x1, x2, x3);
        kzg_ld_chunk(base, cq - 4, lowbits, y0, y1, y2, y3);
        valid = 15;
        ci = cq - 5;
        nw = (nwi & 3) + 1;
        for (i32 k = nw; k < 4; k++) { c3 = c2; c2 = c1; c1 = c0; c0 = 0; }
        left = (i64)T;
        return true;
    }
    KZG_HD void shift_xy() { x0 = y0; x1 = y1; x2 = y2; x3 = y3; valid = (valid & 3u) | ((valid >> 1) & 4u); }
    // Warp-uniform point: promote loaded chunks into the private slots, then refill the load targets.
    KZG_HD void advance() {
        if (!(valid & 1u) && (valid & 4u)) { a0 = x0; a1 = x1; a2 = x2; a3 = x3; shift_xy(); valid |= 1u; }
        if (!(valid & 2u) && (valid & 4u)) { b0 = x0; b1 = x1; b2 = x2; b3 = x3; shift_xy(); valid |= 2u; }
        if (!(valid & 4u)) { kzg_ld_chunk(base, ci, lowbits, x0, x1, x2, x3); ci--; valid |= 4u; }
        if (!(valid & 8u)) { kzg_ld_chunk(base, ci, lowbits, y0, y1, y2, y3); ci--; valid |= 8u; }
    }
    KZG_HD void refill() {     // requires wbits <= 32
        const u32 lo = c3;
        c3 = c2; c2 = c1; c1 = c0;
        if (--nw == 0) {
            if (valid & 1u) {
                c0 = a0; c1 = a1; c2 = a2; c3 = a3;
                a0 = b0; a1 = b1; a2 = b2; a3 = b3;
                valid = (valid & 12u) | ((valid >> 1) & 1u);
            } else if (valid & 4u) {          // ran dry between uniform points (rare): read a load target
                c0 = x0; c1 = x1; c2 = x2; c3 = x3;
                shift_xy();
            } else {
                kzg_ld_chunk(base, ci, lowbits, c0, c1, c2, c3);
                ci--;
            }
            nw = 4;
        }
        const u32 piece = kzg_funnel_r(lo, hi, sh);
        hi = lo;
        win |= (u64)piece << (32 - wbits);
        wbits += 32;
    }
    KZG_HD void ensure32() { if (wbits <= 32) refill(); }
    KZG_HD u32 peek(u32 n) const { return (u32)((win >> 1) >> (63 - n)); }
    KZG_HD void skip(u32 n) { win <<= n; wbits -= (i32)n; left -= (i64)n; }
    KZG_HD u32 read(u32 n) { const u32 v = peek(n); skip(n); return v; }
};

// Forward (LSB-first) bit read for FSE table descriptions; bits past `size` read as zero.
KZG_HD u32 kzg_fwd_bits(const u8* p, u32 size, u32 bitpos, u32 n) {   // n <= 24
    const u32 byte = bitpos >> 3;
    u32 v = 0;
    for (u32 k = 0; k < 4; k++)
        if (byte + k < size) v |= (u32)p[byte + k] << (8 * k);
    return (v >> (bitpos & 7)) & ((1u << n) - 1u);
}
