// Execute stage (s2 + s5) for sm_100a: raw / RLE blocks and ZSTD_execSequence, one warp per frame.
// Device-only code.
//
// A warp walks its frame's blocks in order and, inside a compressed block, the sequence records in
// batches of 32 (one lane per sequence).  Output is assembled in a per-warp shared-memory window and
// leaves for HBM only as 16-byte vector stores, 512 contiguous bytes per warp instruction:
//   * the window covers output positions [T0, T0 + KZG_EX_TILE) of the current block; after every batch
//     the completed 16-byte chunks are flushed, and when the window is full it slides, keeping the last
//     KZG_EX_KEEP bytes as near history;
//   * a match whose source is still inside the window is copied shared -> shared; older history (earlier
//     in the block, earlier blocks of the frame, the dictionary) is read back from global memory / L2,
//     which this warp itself wrote and flushed earlier;
//   * copies are word-granular (aligned 32-bit loads + funnel shift), bytes only at the ragged ends;
//   * matches whose source overlaps output produced by earlier lanes of the same batch wait for those
//     lanes: the batch is resolved in rounds, each round completing every lane whose source lies below
//     the first still-pending lane (usually one round);
//   * literal runs or matches longer than KZG_EX_LONG bytes are done by the whole warp, global -> global.
// No inter-warp synchronisation, no spinning.  What libzstd does here is ZSTD_execSequence, reached from
// the reference through ZSTD_decompressStream (kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:178).
#pragma once
#include "kzg_stages.cuh"

#define KZG_EX_WARPS 4
#define KZG_EX_TILE 8192u
#define KZG_EX_KEEP 4032u           // multiple of 16; TILE - KEEP - 16 >= 32 * 2 * KZG_EX_LONG
#define KZG_EX_LONG 64u
#define KZG_EX_SMEM (KZG_EX_WARPS * KZG_EX_TILE)

namespace kzg_ex {

__device__ __forceinline__ u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ u32 lds8(u32 a) { u32 v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(u32 a, u32 v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lds128(u32 a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(u32 a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// History written earlier by this warp: read at L2 (the coherence point), never from a stale L1 line.
__device__ __forceinline__ u32 ldg32_hist(const u32* p) { return __ldcg(p); }
__device__ __forceinline__ u32 ldg8_hist(const u8* p) { return (u32)__ldcg(p); }

// Source of a per-lane copy into the window.
enum : u32 { SRC_TILE = 0, SRC_GLOBAL = 1, SRC_CONST = 2, SRC_FILL = 3 };

// One lane copies n (<= KZG_EX_LONG) bytes to shared address `d`.  kind selects where the bytes come from:
//   SRC_TILE   shared address `sa` (window -> window, source at least 8 bytes behind the destination)
//   SRC_GLOBAL history at `gp`, re-read from L2          SRC_CONST  read-only data at `gp` (literals)
//   SRC_FILL   the byte `fill`
// Every lane of the warp calls this together (n = 0 for idle lanes): three uniform loops, no syncs.
__device__ __forceinline__ void lane_copy(u32 d, u32 n, u32 kind, u32 sa, const u8* gp, u32 fill) {
    // head: bytes until the destination is word aligned
    u32 h = min(n, (4u - (d & 3u)) & 3u);
    n -= h;
    while (__any_sync(0xFFFFFFFFu, h != 0)) {
        if (h) {
            u32 v;
            if (kind == SRC_TILE) v = lds8(sa);
            else if (kind == SRC_GLOBAL) v = ldg8_hist(gp);
            else if (kind == SRC_CONST) v = __ldg(gp);
            else v = fill;
            sts8(d, v);
            d++; sa++; gp++; h--;
        }
    }
    // words
    u32 nw = n >> 2;
    n &= 3u;
    u32 r = 0, lo = 0;
    const u32* gw = nullptr;
    if (kind == SRC_TILE) { r = (sa & 3u) * 8u; sa &= ~3u; if (nw) lo = lds32(sa); }
    else if (kind == SRC_FILL) { lo = fill * 0x01010101u; }
    else {
        r = (u32)((uintptr_t)gp & 3u) * 8u;
        gw = reinterpret_cast<const u32*>(gp - ((uintptr_t)gp & 3u));
        if (nw) lo = kind == SRC_GLOBAL ? ldg32_hist(gw) : __ldg(gw);
    }
    gp += nw << 2;
    while (__any_sync(0xFFFFFFFFu, nw != 0)) {
        if (nw) {
            // `lo` is the aligned source word holding the first byte still to copy; the next aligned word is
            // fetched only if bytes of it are needed (r != 0) or another destination word follows
            u32 hi = 0;
            const bool more = r != 0 || nw > 1;
            if (kind == SRC_TILE) { sa += 4; if (more) hi = lds32(sa); }
            else if (kind == SRC_FILL) hi = lo;
            else { gw++; if (more) hi = kind == SRC_GLOBAL ? ldg32_hist(gw) : __ldg(gw); }
            sts32(d, __funnelshift_r(lo, hi, r));
            lo = hi;
            d += 4; nw--;
        }
    }
    // tail bytes
    if (kind == SRC_TILE) sa += (r >> 3);           // back to a byte address
    while (__any_sync(0xFFFFFFFFu, n != 0)) {
        if (n) {
            u32 v;
            if (kind == SRC_TILE) v = lds8(sa);
            else if (kind == SRC_GLOBAL) v = ldg8_hist(gp);
            else if (kind == SRC_CONST) v = __ldg(gp);
            else v = fill;
            sts8(d, v);
            d++; sa++; gp++; n--;
        }
    }
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
    if ((uintptr_t)(w.out + w.F) & 15) {      // unaligned start (block start, or after a warp-wide copy)
        const i32 nh = min(upto - w.F, (i32)(16 - ((uintptr_t)(w.out + w.F) & 15)));
        if (!all && w.F + nh > upto) return;
        if ((i32)lane < nh) w.out[w.F + (i32)lane] = (u8)lds8(w.tile + (u32)(w.F - w.T0) + lane);
        w.F += nh;
    }
    const i32 nvec = (upto - w.F) >> 4;
    if (nvec > 0) {
        uint4* g = reinterpret_cast<uint4*>(w.out + w.F);
        const u32 t = w.tile + (u32)(w.F - w.T0);
        for (i32 v = (i32)lane; v < nvec; v += 32) g[v] = lds128(t + ((u32)v << 4));
        w.F += nvec << 4;
    }
    if (all && w.F < upto) {
        const i32 nt = upto - w.F;              // < 16
        if ((i32)lane < nt) w.out[w.F + (i32)lane] = (u8)lds8(w.tile + (u32)(w.F - w.T0) + lane);
        w.F = upto;
    }
}

// Makes room so that positions up to `need_end` fit: slides the window down, keeping KZG_EX_KEEP bytes of
// flushed history.  Requires everything below `cur` (rounded down to 16) to be flushed already.
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
    w.T0 = nT0;
    if (w.Tv < nT0) w.Tv = nT0;
}

}  // namespace kzg_ex

// Executes one compressed block with sequences. Returns true (uniformly) if the block is corrupt.
__device__ __forceinline__ bool kzg_exec_sequences(const KzgBatch& B, const KzgBlock& blk, const KzgCBlock* cb, u8* out, u32 tile, u32 lane) {
    using namespace kzg_ex;
    const u32 nseq = cb->n_seq, lit_regen = cb->lit_regen, lit_type = cb->lit_type;
    const u8* lit = lit_type >= 2 ? B.lit_pool + cb->lit_out : B.src + cb->src + cb->lit_src;
    const u32 rle = lit_type == 1 ? (u32)lit[0] : 0u;
    const u32 lkind = lit_type == 1 ? SRC_FILL : SRC_CONST;
    const u64* seqs = B.seq_pool + cb->seq_out;
    const u32 ri0 = cb->rep_init[0], ri1 = cb->rep_init[1], ri2 = cb->rep_init[2];
    const u64 hist_before_block = blk.out_off + B.dict_size;
    Win w;
    w.tile = tile; w.out = out;
    w.T0 = -(i32)((uintptr_t)out & 15);
    w.Tv = 0; w.F = 0;
    w.boff = (i64)blk.out_off;
    w.dict_end = B.dict + B.dict_size;
    bool err = false;
    u64 carry = 0;                       // last record of the previous batch
    u64 rec_next = lane < nseq ? seqs[lane] : 0ull;
    i32 cur = 0;

    for (u32 b0 = 0; b0 < nseq; b0 += 32) {
        const u32 i = b0 + lane;
        const bool valid = i < nseq;
        const u64 rec = rec_next;
        rec_next = (i + 32 < nseq) ? seqs[i + 32] : 0ull;          // prefetch the next batch
        u64 prev = __shfl_up_sync(0xFFFFFFFFu, rec, 1);
        if (lane == 0) prev = carry;
        const u32 nvalid = min(32u, nseq - b0);
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
        const bool bad = valid && (off == 0 || off >= KZG_OFF_INVALID || (u64)off > hist_before_block + mdst ||
                                   lit_end < p_lit_end || out_end < mdst);
        if (__any_sync(0xFFFFFFFFu, bad)) { err = true; break; }
        const u32 longm = __ballot_sync(0xFFFFFFFFu, valid && (ll > KZG_EX_LONG || ml > KZG_EX_LONG));
        const u32 validm = nvalid == 32 ? 0xFFFFFFFFu : ((1u << nvalid) - 1u);

        u32 todo = validm;                               // lanes not yet executed, processed in order
        while (todo) {
            // next group: the run of lanes before the first long one
            const u32 first = __ffs(todo) - 1;
            const u32 longs = longm & todo;
            const u32 stop = longs ? (u32)__ffs(longs) - 1 : 32u;      // lane index of the next long sequence
            const u32 grp = stop >= 32 ? todo : (todo & ((1u << stop) - 1u));
            if (grp) {
                const u32 lastl = 31 - __clz(grp);
                const i32 ge = (i32)__shfl_sync(0xFFFFFFFFu, out_end, lastl);
                if ((u32)(ge - w.T0) > KZG_EX_TILE) slide(w, cur, lane);
                const bool mine = (grp >> lane) & 1u;
                // literal runs of the group (no dependencies)
                lane_copy(w.tile + (u32)((i32)p_out_end - w.T0), mine ? ll : 0u, lkind, 0u, lit + p_lit_end, rle);
                __syncwarp();
                // matches, in dependency rounds
                bool pending = mine && ml > 0;
                const i32 spos = (i32)mdst - (i32)off;
                // source bytes this match needs to be final: everything below `need`
                const i32 need = off >= ml ? spos + (i32)ml : (i32)mdst;
                for (;;) {
                    const u32 pm = __ballot_sync(0xFFFFFFFFu, pending);
                    if (!pm) break;
                    const i32 fin = (i32)__shfl_sync(0xFFFFFFFFu, mdst, __ffs(pm) - 1);   // all below is final
                    const bool go = pending && need <= fin;
                    // classify the source
                    u32 kind = SRC_TILE, n = 0, sa = 0;
                    const u8* gp = nullptr;
                    bool slow = false;
                    if (go) {
                        n = ml;
                        if (spos >= w.Tv) {
                            if (off >= 8) sa = w.tile + (u32)(spos - w.T0);
                            else slow = true;
                        } else if (spos + (i32)ml <= w.Tv && w.boff + spos >= 0) {
                            kind = SRC_GLOBAL; gp = w.out + spos;
                        } else slow = true;
                        if (slow) n = 0;
                    }
                    lane_copy(w.tile + (u32)((i32)mdst - w.T0), n, kind, sa, gp, 0u);
                    if (__any_sync(0xFFFFFFFFu, slow)) {
                        if (slow) {     // short period, or a source straddling the window / dictionary boundary: bytes in order
                            const u32 d = w.tile + (u32)((i32)mdst - w.T0);
                            for (u32 k = 0; k < ml; k++) {
                                const i32 p = spos + (i32)k;
                                const u32 v = p >= w.Tv ? lds8(w.tile + (u32)(p - w.T0)) : hist_byte(w, p);
                                sts8(d + k, v);
                            }
                        }
                    }
                    if (go) pending = false;
                    __syncwarp();
                }
                cur = ge;
                flush(w, cur, false, lane);
                todo &= ~grp;
            }
            if (stop < 32) {
                // long sequence `stop`: the warp copies it straight to global memory
                const u32 l_ll = __shfl_sync(0xFFFFFFFFu, ll, stop), l_ml = __shfl_sync(0xFFFFFFFFu, ml, stop);
                const u32 l_ple = __shfl_sync(0xFFFFFFFFu, p_lit_end, stop), l_poe = __shfl_sync(0xFFFFFFFFu, p_out_end, stop);
                const u32 l_off = __shfl_sync(0xFFFFFFFFu, off, stop);
                flush(w, cur, true, lane);
                __syncwarp();
                u8* d = out + l_poe;
                if (lit_type == 1) warp_set_global(d, rle, l_ll, lane);
                else warp_copy_global<false>(d, lit + l_ple, l_ll, lane);
                __threadfence_block();
                __syncwarp();
                d += l_ll;
                const i32 sp = (i32)(l_poe + l_ll) - (i32)l_off;
                if (l_off >= l_ml && w.boff + sp >= 0) {
                    warp_copy_global<true>(d, out + sp, l_ml, lane);      // disjoint source inside the frame
                } else {
                    // overlapping (period l_off) or reaching into the dictionary: every byte equals one of the
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
                todo &= ~(1u << stop);
            }
            (void)first;
        }
    }
    if (err) return true;
    // tail: what is left in the window, then the literals after the last sequence
    flush(w, cur, true, lane);
    __syncwarp();
    {
        const u32 used = (u32)(carry >> 28) & 0x3FFFFu;
        if (lit_regen > used) {
            if (lit_type == 1) warp_set_global(out + cur, rle, lit_regen - used, lane);
            else warp_copy_global<false>(out + cur, lit + used, lit_regen - used, lane);
        }
    }
    __threadfence_block();
    __syncwarp();
    return false;
}

__global__ void __launch_bounds__(KZG_EX_WARPS * 32) kzg_exec_kernel(KzgBatch B) {
    using namespace kzg_ex;
    extern __shared__ uint4 smem4[];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 tile = (u32)__cvta_generic_to_shared(smem4) + warp * KZG_EX_TILE;
    if (B.ctr->overflow & 3u) return;
    for (;;) {
        u32 f = 0;
        if (lane == 0) f = atomicAdd(&B.ctr->ticket_exec, 1u);
        f = __shfl_sync(0xFFFFFFFFu, f, 0);
        if (f >= B.n_frames) break;
        const KzgFrame* fr = &B.frames[f];
        if (fr->status != KZG_OK) continue;
        u8* const fdst = B.dst + B.dst_off[f];
        const u32 nblk = fr->n_blocks, first = fr->first_block;
        bool failed = false;
        for (u32 k = 0; k < nblk && !failed; k++) {
            const KzgBlock blk = B.blocks[first + k];
            u8* out = fdst + blk.out_off;
            if (blk.type == 0) warp_copy_global<false>(out, B.src + blk.src, blk.regen, lane);
            else if (blk.type == 1) warp_set_global(out, B.src[blk.src], blk.regen, lane);
            else {
                const KzgCBlock* cb = &B.cblocks[blk.cidx];
                if (cb->n_seq == 0) {
                    const u8* lit = cb->lit_type >= 2 ? B.lit_pool + cb->lit_out : B.src + cb->src + cb->lit_src;
                    if (cb->lit_type == 1) warp_set_global(out, lit[0], blk.regen, lane);
                    else warp_copy_global<false>(out, lit, blk.regen, lane);
                } else failed = kzg_exec_sequences(B, blk, cb, out, tile, lane);
            }
            __threadfence_block();
            __syncwarp();
        }
        if (failed && lane == 0) B.frames[f].status = KZG_E_CORRUPTION;
    }
}
