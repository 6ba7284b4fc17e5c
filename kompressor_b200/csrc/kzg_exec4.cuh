// Execute stage (s5, ZSTD_execSequence) for sm_100a, CTA-per-frame variant: thirty-two warps share ONE frame.
//
// The warp-per-frame kernel (kzg_exec.cuh) needs thousands of frames in flight; a batch of sixteen 4 MiB frames
// (BASELINE configs[4]: level 19, 8 MiB window) or a single frame behind the streaming shim leaves the GPU idle while
// one warp walks 140 000 sequences.  Here a CTA takes a frame, and inside a block it executes 1 024 sequences per step,
// one lane per sequence:
//   * the window is the WHOLE block (128 KiB + alignment) in shared memory: every match source inside the block is
//     shared -> shared, nothing slides, and the block leaves as one coalesced 16-byte-vector flush;
//   * sources before the block (earlier blocks of the frame, the dictionary) are read from global memory / L2 by the
//     lane itself -- with thirty-two warps per SM the load latency is covered by the other warps; at level 19 these are
//     the majority (median distance 120 KB) and have no dependencies inside the step;
//   * the literals of a step are contiguous in the literal stream: the CTA stages them with 16-byte loads, lanes copy
//     them shared -> shared;
//   * matches whose source lies inside the step's own output wait for exactly the sequences that produce it: a lane
//     finds them once by two binary searches over the step's output ends (shared memory), every warp publishes a
//     32-bit word of completed lanes, and a pending lane goes as soon as the bits of its producers are set.  Warps
//     poll; dependencies only point backwards, so the earliest pending match can always go and nothing deadlocks.
//     No CTA barrier is needed until the step ends (a frontier rule with a barrier per round needed 54 rounds per
//     step on level-19 text, granule counters 18);
//   * long literal runs and matches (>= 256 bytes) are copied by their whole warp.
// Raw / RLE / literal-only blocks were written by the bulk kernel and are skipped.  This kernel executes the frames the
// hot kernel left on the deferred list (more than 8 blocks, long runs) and, for small batches, every frame.
// What libzstd does here is ZSTD_execSequence, reached from the reference through ZSTD_decompressStream
// (kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:178).
#pragma once
#include "kzg_exec.cuh"
#ifdef KZG_X4_TIMING
#include <stdio.h>
#endif

#ifndef KZG_X4_WARPS
#define KZG_X4_WARPS 32                                 // measured on 16 / 64 level-19 frames of 4 MiB: 8 warps 7.0 / 7.3 ms, 12: 5.8 / 6.1, 16: 5.1 / 5.5, 20: 4.7 / 5.0, 24: 4.5 / 4.8, 32: 4.3 / 4.6
#endif
#define KZG_X4_THREADS (KZG_X4_WARPS * 32)
#define KZG_X4_WIN (KZG_BLOCK_MAX + 48u)                // the block + alignment + the words lane copies may read past an end
#ifndef KZG_X4_LITCAP
#define KZG_X4_LITCAP 12288u                            // most literal bytes of one step that are staged (with 32 warps the per-lane match staging takes 80 KB)
#endif
#define KZG_X4_LIT_BYTES (KZG_X4_LITCAP + 48u)
#define KZG_X4_MISC 256u                                // first-pending positions (2 x 16), literal range, flags
#define KZG_X4_MSTAGE 80u                               // per-lane staging of a match source read from global memory (<= 64 bytes)
#define KZG_X4_GLONG 64u                                // longer ones are copied by the whole warp
#define KZG_X4_OEND_BYTES (KZG_X4_THREADS * 4u)         // where every sequence of the step ends
#define KZG_X4_SMEM (KZG_X4_WIN + KZG_X4_LIT_BYTES + KZG_X4_MISC + KZG_X4_THREADS * KZG_X4_MSTAGE + KZG_X4_OEND_BYTES)
#define KZG_X4_COOP 256u                                // literal runs / matches of at least this many bytes: the whole warp copies
#define KZG_X4_INF 0x7FFFFFFFu

namespace kzg_x4 {
using namespace kzg_ex;

// The warp copies n bytes from global memory (earlier output of the frame, visible at L2; negative frame positions: the
// dictionary) to shared address d: 128 bytes per trip, four independent loads per lane.
__device__ __forceinline__ void warp_copy_hist(u32 d, u32 n, const u8* out, i32 sp, i64 boff, const u8* dict_end, u32 lane) {
    for (u32 k0 = 0; k0 < n; k0 += 128) {
        u32 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const u32 k = k0 + lane + 32u * j;
            const i64 fp = boff + sp + (i64)k;
            v[j] = k < n ? (fp >= 0 ? (u32)__ldcg(out + sp + (i32)k) : (u32)dict_end[fp]) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const u32 k = k0 + lane + 32u * j;
            if (k < n) sts8(d + k, v[j]);
        }
    }
}

// One lane fills n bytes at shared address d with the period of p (1..7) bytes that ends right in front of d (an
// overlapping match with a short distance).  The period is read once; every output word is a byte permutation of it, so
// nothing waits on the lane's own stores (byte by byte, every load would wait for the store before it).
__device__ __forceinline__ void lane_fill_period(u32 d, u32 n, u32 p) {
    if (n == 0) return;
    const u32 sa = d - p, A = sa & ~3u, r = (sa & 3u) << 3;
    const u32 w0 = lds32(A), w1 = lds32(A + 4), w2 = lds32(A + 8);
    const u32 s0 = __funnelshift_r(w0, w1, r), s1 = __funnelshift_r(w1, w2, r);     // bytes sa .. sa + 7; the first p are the period
    u32 m = 0;                                      // position inside the period of the next output byte
    auto next4 = [&]() -> u32 {                     // the next four output bytes
        u32 sel = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) { sel |= m << (4 * j); m = m + 1 == p ? 0u : m + 1; }
        return __byte_perm(s0, s1, sel);
    };
    const u32 h = min(n, (0u - d) & 3u);
    if (h) {
        const u32 save = m;
        const u32 v = next4();
        sts8(d, v);
        if (h > 1) sts8(d + 1, v >> 8);
        if (h > 2) sts8(d + 2, v >> 16);
        m = save + h; while (m >= p) m -= p;
        d += h; n -= h;
    }
    for (u32 rem = n >> 2; rem; rem--) { sts32(d, next4()); d += 4; }
    const u32 t = n & 3u;
    if (t) {
        const u32 v = next4();
        sts8(d, v);
        if (t > 1) sts8(d + 1, v >> 8);
        if (t > 2) sts8(d + 2, v >> 16);
    }
}

}  // namespace kzg_x4

// Executes one compressed block with sequences into the shared-memory window and flushes it.  Returns (uniformly over
// the CTA) 0 = done, 1 = the block is corrupt.
__device__ __forceinline__ u32 kzg_exec4_block(const KzgBatch& B, const KzgBlock& blk, const KzgCBlock* cb, u8* out, u32 smem) {
    using namespace kzg_x4;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32 nseq = cb->n_seq, lit_regen = cb->lit_regen, lit_type = cb->lit_type, regen = blk.regen;
    const u8* lit = lit_type >= 2 ? B.lit_pool + cb->lit_out : B.src + cb->src + cb->lit_src;
    const bool lfill = lit_type == 1;
    const u32 rle = lfill ? (u32)__ldg(lit) : 0u;
    const u64* seqs = B.seq_pool + cb->seq_out;
    const u32 ri0 = cb->rep_init[0], ri1 = cb->rep_init[1], ri2 = cb->rep_init[2];
    const u64 hist64 = blk.out_off + B.dict_size;
    const u32 hist_before_block = hist64 > 0x40000000ull ? 0x40000000u : (u32)hist64;
    const i64 boff = (i64)blk.out_off;
    const u8* dict_end = B.dict + B.dict_size;
    const u32 sh = (u32)((uintptr_t)out & 15);
    const u32 win = smem + sh;                          // shared address of block position 0
    const u32 lstage = smem + KZG_X4_WIN;
    volatile u32* misc = reinterpret_cast<volatile u32*>(__cvta_shared_to_generic(smem + KZG_X4_WIN + KZG_X4_LIT_BYTES));
    // misc[0..15]: completed lanes of every warp; misc[32], misc[33]: literal range of the step; misc[34]: its first output
    volatile u32* oend = reinterpret_cast<volatile u32*>(__cvta_shared_to_generic(smem + KZG_X4_WIN + KZG_X4_LIT_BYTES + KZG_X4_MISC + KZG_X4_THREADS * KZG_X4_MSTAGE));

    u64 rec_n = tid < nseq ? __ldg(seqs + tid) : 0ull, prev_n = (tid && tid < nseq) ? __ldg(seqs + tid - 1) : 0ull;         // the step's records are fetched one step ahead
#ifdef KZG_X4_TIMING
    long long tA = 0, tB = 0, tC = 0, tD = 0; u32 nsteps = 0; long long tripsum = 0;
#endif
    for (u32 sb0 = 0; sb0 < nseq; sb0 += KZG_X4_THREADS) {
#ifdef KZG_X4_TIMING
        const long long c0 = clock64();
#endif
        const u32 i = sb0 + tid;
        const bool valid = i < nseq;
        const u64 rec = rec_n, prev = prev_n;
        {
            const u32 in = i + KZG_X4_THREADS;
            rec_n = in < nseq ? __ldg(seqs + in) : 0ull;
            prev_n = in < nseq ? __ldg(seqs + in - 1) : 0ull;
        }
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
        if (tid == 0) { misc[32] = p_lit_end; misc[34] = p_out_end; misc[35] = 0u; }
        if (lane == 0) misc[warp] = 0u;
        oend[tid] = valid ? out_end : 0x7FFFFFFFu;
        if (valid && (i + 1 == nseq || tid == KZG_X4_THREADS - 1)) misc[33] = lit_end;
        if (__syncthreads_or(bad)) return 1u;

        // ---- matches whose source lies before the block (earlier blocks, the dictionary): no dependencies inside the block.
        // The 16-byte chunks covering a short source inside the frame are requested now and fly during the literal pass.
        const i32 spos = (i32)mdst - (i32)off;
        const bool isg = ml > 0 && spos + (i32)ml <= 0;                 // entirely before the block
        const i64 fpos = boff + spos;                                    // frame position of the source (negative: dictionary)
        const bool gstage = isg && fpos >= 0 && ml <= KZG_X4_GLONG;
        const u32 mstage = smem + KZG_X4_WIN + KZG_X4_LIT_BYTES + KZG_X4_MISC + tid * KZG_X4_MSTAGE;
        u32 ga;
        {
            const u8* g = out + spos;
            ga = (u32)((uintptr_t)g & 15);
            const u32 nch = gstage ? (ga + ml + 15u) >> 4 : 0u;         // <= 5
            g -= ga;
            cp_async16_if(nch > 0, mstage, g);
            cp_async16_if(nch > 1, mstage + 16, g + 16);
            cp_async16_if(nch > 2, mstage + 32, g + 32);
            cp_async16_if(nch > 3, mstage + 48, g + 48);
            cp_async16_if(nch > 4, mstage + 64, g + 64);
            cp_commit();
        }

        // ---- literals of this step: staged by the CTA, then one lane per run
        const u32 l0 = misc[32], l1 = misc[33];
        // (a step with more literals than the staging area holds stages the first KZG_X4_LITCAP bytes; the runs beyond come
        //  straight from global memory)
        const u32 l1s = min(l1, l0 + KZG_X4_LITCAP);
        const bool staged = !lfill && l1 >= l0;
        const u32 la = (u32)((uintptr_t)(lit + l0) & 15);
        if (staged) {
            const uint4* g4 = reinterpret_cast<const uint4*>(lit + l0 - la);
            const u32 nch = (la + (l1s - l0) + 15u) >> 4;
            for (u32 c = tid; c < nch; c += KZG_X4_THREADS) sts128(lstage + (c << 4), __ldg(g4 + c));
        }
        __syncthreads();
        {
            const u32 d = win + p_out_end;
            const u32 n = ll < KZG_X4_COOP ? ll : 0u;
            if (lfill || (staged && lit_end <= l1s)) lane_copy(d, n, lstage + la + (p_lit_end - l0), lfill, rle);
            else for (u32 k = 0; k < n; k++) sts8(d + k, (u32)__ldg(lit + p_lit_end + k));
            u32 longl = __ballot_sync(0xFFFFFFFFu, ll >= KZG_X4_COOP);
            while (longl) {
                const u32 s = __ffs(longl) - 1;
                longl &= longl - 1;
                const u32 dd = win + __shfl_sync(0xFFFFFFFFu, p_out_end, s), nn = __shfl_sync(0xFFFFFFFFu, ll, s);
                const u8* g = lit + __shfl_sync(0xFFFFFFFFu, p_lit_end, s);
                for (u32 k = lane; k < nn; k += 32) sts8(dd + k, lfill ? rle : (u32)__ldg(g + k));
            }
        }

        // ---- matches whose source lies before the block: entirely inside the frame and short, the lane's 16-byte chunks were
        // requested above and are copied from its staging area now; otherwise the whole warp copies
        cp_wait<0>();
        __syncwarp();
        lane_copy(win + mdst, gstage ? ml : 0u, mstage + ga, false, 0u);
        {
            u32 longm = __ballot_sync(0xFFFFFFFFu, isg && !gstage);
            while (longm) {
                const u32 s = __ffs(longm) - 1;
                longm &= longm - 1;
                warp_copy_hist(win + __shfl_sync(0xFFFFFFFFu, mdst, s), __shfl_sync(0xFFFFFFFFu, ml, s), out,
                               __shfl_sync(0xFFFFFFFFu, spos, s), boff, dict_end, lane);
            }
        }
        // source bytes this match needs to be final: everything below `need`
        const i32 need = off >= ml ? spos + (i32)ml : (i32)mdst;
        const bool inwin = spos >= 0;                                    // else it straddles the start of the block
        const bool islong = ml >= KZG_X4_COOP;
        // a source that ends below this step's first output is final already
        const u32 S0 = misc[34];
        const bool early = ml > 0 && !isg && need <= (i32)S0;
        bool pending = ml > 0 && !isg && !early;
        // the sequences of this step whose output overlaps the source: ja .. jb (producers are always earlier sequences)
        u32 ja = 0, jb = 0;
        if (pending) {
            const u32 a = (u32)max(spos, (i32)S0), b = (u32)need;         // source range inside the step's output: [a, b)
            u32 lo = 0, hi = tid;                                       // first j with oend[j] > a
            while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (oend[mid] > a) hi = mid; else lo = mid + 1; }
            ja = lo;
            hi = tid;                                                   // first j with oend[j] >= b (its output starts below b)
            while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (oend[mid] >= b) hi = mid; else lo = mid + 1; }
            jb = min(lo, tid - 1u);                                     // (tid >= 1 here: sequence 0 of a step has no producer in it)
            if (tid == 0 || ja > jb) { ja = 1; jb = 0; }                // no producers: the source is literals of this very sequence
        }
        // the words of completed lanes that hold the producers' bits, and the bits
        const u32 w0 = ja >> 5, w1 = jb >> 5;
        const u32 mask0 = ja > jb ? 0u : ((0xFFFFFFFFu << (ja & 31u)) & (w1 == w0 ? 0xFFFFFFFFu >> (31u - (jb & 31u)) : 0xFFFFFFFFu));
        const u32 mask1 = 0xFFFFFFFFu >> (31u - (jb & 31u));
#ifdef KZG_X4_TIMING
        const long long c1 = clock64();
#endif
        u32 donem = 0;                          // completed lanes of this warp (what misc[warp] holds)
        u32 trips = 0;
        bool go = early;                        // first trip: the matches whose sources are final already
        bool fin = !pending && !early;          // lanes with nothing (left) to copy
        for (;;) {
            const u32 gm = __ballot_sync(0xFFFFFFFFu, go);
            if (gm) {
                const bool fast = inwin && off >= 8u && ml < KZG_X4_COOP;
                const bool period = inwin && off < 8u && ml < KZG_X4_COOP;
                lane_copy(win + mdst, (go && fast) ? ml : 0u, win + (u32)spos, false, 0u);
                if (__any_sync(0xFFFFFFFFu, go && period)) lane_fill_period(win + mdst, (go && period) ? ml : 0u, off);
                if (go && !fast && !period && ml < KZG_X4_COOP) {       // a source reaching back before the block: bytes in order
                    for (u32 k = 0; k < ml; k++) {
                        const i32 p = spos + (i32)k;
                        u32 v;
                        if (p >= 0) v = lds8(win + (u32)p);
                        else { const i64 fp = boff + p; v = fp >= 0 ? (u32)__ldcg(out + p) : (u32)dict_end[fp]; }
                        sts8(win + mdst + k, v);
                    }
                }
                u32 longm = __ballot_sync(0xFFFFFFFFu, go && ml >= KZG_X4_COOP);
                while (longm) {
                    // a long match inside the window: the warp copies min(distance, 1 KiB) bytes per step; the distance doubles
                    // while the period unrolls (runs)
                    const u32 s = __ffs(longm) - 1;
                    longm &= longm - 1;
                    const u32 dm = __shfl_sync(0xFFFFFFFFu, mdst, s), nn = __shfl_sync(0xFFFFFFFFu, ml, s);
                    const i32 sp = __shfl_sync(0xFFFFFFFFu, spos, s);
                    u32 dist = __shfl_sync(0xFFFFFFFFu, off, s);
                    for (u32 done = 0; done < nn;) {
                        const u32 n = min(min(nn - done, dist), 1024u);
                        for (u32 k = lane; k < n; k += 32) {
                            const i32 p = (i32)(dm + done + k) - (i32)dist;
                            u32 v;
                            if (p >= 0) v = lds8(win + (u32)p);
                            else { const i64 fp = boff + p; v = fp >= 0 ? (u32)__ldcg(out + p) : (u32)dict_end[fp]; }
                            sts8(win + dm + done + k, v);
                        }
                        __syncwarp();
                        done += n;
                        if (dist < 1024u && (i32)(dm + done) - sp >= (i32)(2u * dist)) dist <<= 1;
                    }
                }
            }
            // ---- publish the lanes that are complete (their bytes are in shared memory first), then see who can go next
            fin = fin || go;
            pending = pending && !go;
            const u32 nowdone = __ballot_sync(0xFFFFFFFFu, fin);
            if (nowdone != donem) {
                asm volatile("fence.acq_rel.cta;" ::: "memory");
                if (lane == 0) misc[warp] = nowdone;
                donem = nowdone;
            }
            if (nowdone == 0xFFFFFFFFu) break;
            // poll until a lane of this warp can go (its producers' bits are set; a producer in this very warp counts at once)
            for (;;) {
                go = false;
                if (pending) {
                    const u32 d0 = w0 == warp ? nowdone : misc[w0];
                    bool ok = (d0 & mask0) == mask0;
                    if (w1 != w0) {
                        const u32 d1 = w1 == warp ? nowdone : misc[w1];
                        ok = ok && (d1 & mask1) == mask1;
                        for (u32 w = w0 + 1; w < w1 && ok; w++) ok = (w == warp ? nowdone : misc[w]) == 0xFFFFFFFFu;
                    }
                    go = ok;
                }
                if (__any_sync(0xFFFFFFFFu, go)) break;
                if (++trips > (1u << 24)) { if (lane == 0) misc[35] = 1u; break; }      // cannot happen; never spin forever on a GPU
            }
            if (trips > (1u << 24)) break;
        }
#ifdef KZG_X4_TIMING
        const long long c2 = clock64();
#endif
        __syncthreads();                        // the step is complete: the next one may read all of it
#ifdef KZG_X4_TIMING
        { const long long c3 = clock64(); tA += c1 - c0; tB += c2 - c1; tC += c3 - c2; nsteps++; tripsum += trips; }
#endif
        if (misc[35]) return 1u;
    }
    // ---- the literals after the last sequence, then the whole block leaves
    {
        const u32 used = cb->lit_used, trail = lit_regen - used, o0 = regen - trail;
        for (u32 k = tid; k < trail; k += KZG_X4_THREADS) sts8(win + o0 + k, lfill ? rle : (u32)__ldg(lit + used + k));
    }
    __syncthreads();
    {
        const u32 head = min(regen, (16u - sh) & 15u);
        if (tid < head) out[tid] = (u8)lds8(win + tid);
        const u32 nvec = (regen - head) >> 4;
        uint4* g = reinterpret_cast<uint4*>(out + head);
        const u32 t = win + head;
        for (u32 v = tid; v < nvec; v += KZG_X4_THREADS) g[v] = lds128(t + (v << 4));
        const u32 donev = head + (nvec << 4);
        if (tid < regen - donev) out[donev + tid] = (u8)lds8(win + donev + tid);
    }
    __threadfence();                    // later blocks read this one back from L2
    __syncthreads();
#ifdef KZG_X4_TIMING
    if (blockIdx.x == 0 && (tid & 31) == 0 && blk.out_off < (1u << 20)) printf("x4 warp %u: steps %u  phaseA %lld  chains %lld  barrier %lld  trips %lld (per step: %lld %lld %lld)\n", tid >> 5, nsteps, tA, tB, tC, tripsum, tA / nsteps, tB / nsteps, tC / nsteps);
    (void)tD;
#endif
    return 0u;
}

// ALL = false: the frames on the deferred list; ALL = true: every frame of the batch (small batches).
template <bool ALL>
__global__ void __launch_bounds__(KZG_X4_THREADS, 1) kzg_exec4_kernel(KzgBatch B) {
    extern __shared__ uint4 smem4[];
    __shared__ u32 s_ticket;
    const u32 smem = (u32)__cvta_generic_to_shared(smem4);
    if (B.ctr->overflow & 3u) return;
    const u32 total = ALL ? B.n_frames : B.ctr->n_defer;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = atomicAdd(&B.ctr->ticket_exec4, 1u);
        __syncthreads();
        u32 f = s_ticket;
        if (f >= total) break;
        if (!ALL) {
            f = B.dlist[f];
            if (!(f & KZG_DEFER_BIG)) continue;         // long runs in a small frame: the first design's full variant
            f &= ~KZG_DEFER_BIG;
        }
        const KzgFrame* fr = &B.frames[f];
        if (fr->status != KZG_OK) continue;
        u8* const fdst = B.dst + B.dst_off[f];
        const u32 nblk = fr->n_blocks, first = fr->first_block;
        bool failed = false;
        u32 kfail = 0;
        for (u32 k = 0; k < nblk && !failed; k++) {
            const KzgBlock blk = B.blocks[first + k];
            // raw / RLE blocks and literal-only blocks were written by kzg_bulk_kernel
            if (blk.type == 2) {
                const KzgCBlock* cb = &B.cblocks[blk.cidx];
                if (cb->n_seq) { failed = kzg_exec4_block(B, blk, cb, fdst + blk.out_off, smem) == 1u; kfail = k; }
            }
        }
        if (failed && threadIdx.x == 0) { B.frames[f].status = KZG_E_CORRUPTION; B.frames[f].fail_block = kfail; }
    }
}
