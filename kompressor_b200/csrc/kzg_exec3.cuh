// Execute stage (s5, ZSTD_execSequence) for sm_100a, third design: one THREAD per frame, 32 frames per warp.
//
// Executing sequences is serial inside a frame (a match may read what the previous sequence wrote), so the two
// earlier designs (kzg_exec.cuh: a lane per sequence, dependency rounds; kzg_exec2.cuh: a lane per output word)
// spend most of their instructions on coordination between lanes: ballots, rounds, ragged ends, idle lanes.
// Here nothing is coordinated.  Every lane owns one frame and executes its sequences in order, like a CPU does, and
// the 32 lanes of a warp are 32 independent frames.  What makes that affordable:
//   * the output window of a lane is a ring of W bytes in shared memory, WORD-INTERLEAVED across the warp (word i of
//     lane L at byte i * 128 + L * 4): whatever position a lane is at, it reads and writes bank L, so no access to
//     the rings ever conflicts;
//   * a lane copies in passes of up to 16 bytes: five aligned source words + funnel shifts give sixteen source bytes,
//     a second set of funnel shifts aligns them with the destination, the partial first word is merged with what the
//     ring holds, and four full words are stored -- no byte stores, no read-modify-write beyond one word;
//     overlapping matches copy `distance` bytes per pass and double the distance as the period unrolls;
//   * a lane is its own producer and consumer: a *classifier* step reads the next sequence record, splits it into
//     entries (<= 32 literal bytes, and <= 32 match bytes when the source has left the ring), issues the cp.async
//     (LDGSTS) copies that stage the literals and the far match source in the lane's private staging area, and queues
//     the entry in a small per-lane FIFO in shared memory; the *executor* step pops an entry eight loop rounds after it was
//     queued, when its copies have landed: the hardware tracks cp.async groups per WARP (LDGDEPBAR / DEPBAR on a
//     warp scoreboard), so the warp commits one group per round and waits for the group of eight rounds ago, and
//     every entry carries the round it was queued in.  No load ever has a register destination on the hot path except the record itself (L1-resident,
//     fetched three records ahead);
//   * the rings are flushed by the warp together, transposed: four lanes take one frame's next 64 bytes, so a store
//     instruction writes eight 64-byte runs (full 32-byte sectors); far match sources are read back from there
//     (L2) by cp.async after a block-level fence.
// Blocks the bulk kernel wrote (raw, RLE, literal-only, long literal runs) are holes the executor jumps over.
// Frames with a dictionary are left to the first design (the host selects it for such batches).
// What libzstd does here is ZSTD_execSequence, reached from the reference through ZSTD_decompressStream
// (kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:178).
#pragma once
#include "kzg_stages.cuh"
#include "kzg_simt.cuh"

#define KZG_X3_D 7u                 // most entries queued ahead of the executor
#define KZG_X3_FD 8u                // FIFO slots per lane
#define KZG_X3_AGE 8                // an entry is executed this many loop rounds after it was queued: its copies have landed
#define KZG_X3_PIECE 32u            // most literal bytes / far match bytes of one entry
#define KZG_X3_NEAR_PIECE 128u      // most match bytes of one entry whose source is in the ring
#define KZG_X3_FLUSH 256u           // a lane this far ahead of its flushed position triggers a flush of the warp
#define KZG_X3_LITRING 128u
#define KZG_X3_FARCHUNKS 16u

template <u32 W> struct KzgX3 {
    static_assert(W >= 512 && (W & (W - 1)) == 0, "ring size");
    static constexpr u32 WW = W / 4;                                // ring rows (words per lane)
    static constexpr u32 RING = 0;                                  // (WW + 5) rows x 128 B; rows past WW are scratch
    static constexpr u32 FIFO = (WW + 5) * 128;                     // KZG_X3_FD slots x 32 lanes x 16 B
    static constexpr u32 LIT = FIFO + KZG_X3_FD * 512;              // per lane: 128 B ring + 48 B mirror of its start
    static constexpr u32 LIT_STRIDE = KZG_X3_LITRING + 48;          // 176 = 16 x 11
    static constexpr u32 FAR = LIT + 32 * LIT_STRIDE;               // per lane: 16 chunks of 16 B + 16
    static constexpr u32 FAR_STRIDE = KZG_X3_FARCHUNKS * 16 + 16;   // 272 = 16 x 17
    static constexpr u32 BYTES = FAR + 32 * FAR_STRIDE + 16;
};

namespace kzg_x3 {

enum : u32 { K_SEQ = 0u, K_SKIP = 1u, K_START = 2u, K_END = 3u };   // entry kinds (word 0, bits 30..31)
enum : u32 { ST_FRAME = 0u, ST_BLOCK = 1u, ST_SEQ = 2u, ST_DONE = 3u };
constexpr u32 E_FAR = 1u << 29;                                     // entry flag: the match source is staged
constexpr u32 LEN_MASK = 0x3FFFFu;
constexpr u32 STAMP_MASK = 0x3FFFu;                                 // 14-bit round / chunk counters ride in the entries' spare bits

KZG_SIMT_D u32 mn(u32 a, u32 b) { return a < b ? a : b; }

}  // namespace kzg_x3

// One warp: every lane takes frames by ticket until none are left.  sb: shared address of KzgX3<W>::BYTES.
template <u32 W>
KZG_SIMT_D void kzg_exec3_warp(const KzgBatch& B, u32 sb, u32 lane) {
    using namespace simt;
    using namespace kzg_x3;
    typedef KzgX3<W> L;
    constexpr u32 M = L::WW - 1u;
    constexpr u32 NEAR_MAX = W - 48u;               // a source this close is still in the ring when its match executes
    constexpr u32 LOOK_MAX = W - 384u;              // bytes the classifier may run ahead of the executor
    constexpr u32 D = KZG_X3_D;
    const u32 ringL = sb + L::RING + (lane << 2);
    const u32 fifoL = sb + L::FIFO + (lane << 4);
    const u32 litL = sb + L::LIT + lane * L::LIT_STRIDE;
    const u32 farL = sb + L::FAR + lane * L::FAR_STRIDE;
    const u32 nframes = B.n_frames;

    // ---- executor state.  Positions are in "X space": frame position + (frame destination & 15), so that X and the
    // global address agree modulo 16 and ring row = (X >> 2) & M.
    u32 c = 0, m = 0, e = 0;            // cursor, start of the match part, end of the current entry
    u32 F = 0;                          // positions below F are in global memory
    u32 lit0 = 0, msx = 0, sfix = 0, dist = 1;      // literal byte of position p: lit0 + p; far: msx + p; near: p - dist
    bool isnear = true;
    u8* gal = nullptr;                  // global address of X = 0 (16-byte aligned)
    u32 head = 0, cnt = 0;              // FIFO: next slot to pop, entries queued
    u32 lit_done = 0;                   // literal ring space below this has been consumed
    u32 drain = 0, drain_n = 0;         // K_END (1) / K_SKIP (2) being executed: waits for the warp's flush
    u32 ex_epoch = 0;                   // frames started by the executor
    bool want_flush = false;
    // ---- classifier state
    u32 st = ST_FRAME;
    u32 cl_epoch = 0;                   // frames started by the classifier
    u32 cf = 0, cph = 0, cc = 0;        // frame, its X-space start, X of the next entry
    u32 Tv = 0, hole_lo = 0, hole_hi = 0;   // ring holds nothing below Tv; [hole_lo, hole_hi): newest range written by the bulk kernel
    u8* cgal = nullptr;
    u32 bk = 0, nblk = 0, first = 0;
    const u64* seqs = nullptr;
    u32 nseq = 0, ri = 0;
    u64 rq0 = 0, rq1 = 0, rq2 = 0;      // records ri, ri + 1, ri + 2
    u32 p_lit_end = 0, p_out_end = 0, rep0 = 1, rep1 = 4, rep2 = 8;
    bool have = false, is_tail = false, tail_done = false, bulkl = false, lfill = false;
    u32 rl = 0, rmm = 0, off = 1, tail = 0, fillw = 0;
    const u8* lit16 = nullptr;
    u32 lal = 0, lp = 0, lbase = 0, lw = 0;         // literal stream: alignment, position, ring-space base, staged up to (ring space)
    u32 fw = 0, fdone = 0;                          // far staging: chunks allocated so far / free below (mod 2^14)
    u32 it = 0;                                     // loop round

    auto emit = [&](u32 w0, u32 w1, u32 w2, u32 w3) {
        sts128(fifoL + (((head + cnt) & (KZG_X3_FD - 1u)) << 9), make_uint4(w0, w1 | ((it & STAMP_MASK) << 18), w2, w3));
        cnt++;
    };
    auto hole = [&](u32 n) {            // the bulk kernel wrote [cc, cc + n)
        emit((K_SKIP << 30) | n, 0u, (fw & STAMP_MASK) << 18, 0u);
        hole_lo = cc; hole_hi = cc + n; Tv = hole_hi; cc += n;
    };

    // ---- classifier: frame and block transitions (rare)
    auto transition = [&]() {
        if (st == ST_FRAME) {
            cf = atomic_add(&B.ctr->ticket_exec, 1u);
            if (cf >= nframes) { st = ST_DONE; return; }
            const KzgFrame* fr = &B.frames[cf];
            if (fr->status != KZG_OK) return;
            if (fr->out_size >= 0x7FFF0000ull) {            // positions are 32-bit here: left to the first design's full variant
                B.dlist[atomic_add(&B.ctr->n_defer, 1u)] = cf;
                return;
            }
            u8* fdst = B.dst + B.dst_off[cf];
            cph = (u32)((uintptr_t)fdst & 15);
            cgal = fdst - cph;
            nblk = fr->n_blocks; first = fr->first_block; bk = 0;
            cc = Tv = hole_lo = hole_hi = cph;
            emit((K_START << 30) | cph, 0u, (u32)(uintptr_t)cgal, (u32)((u64)(uintptr_t)cgal >> 32));
            cl_epoch++;
            st = ST_BLOCK;
            return;
        }
        // ST_BLOCK
        if (bk == nblk) { emit((K_END << 30), 0u, 0u, 0u); st = ST_FRAME; return; }
        const KzgBlock* blk = &B.blocks[first + bk];
        const KzgCBlock* cb = blk->type == 2 ? &B.cblocks[blk->cidx] : nullptr;
        if (!cb || cb->n_seq == 0) {
            if (blk->regen) hole(blk->regen);
            bk++;
            return;
        }
        nseq = cb->n_seq; ri = 0;
        seqs = B.seq_pool + cb->seq_out;
        rq0 = ldg64_ro(seqs);
        rq1 = nseq > 1 ? ldg64_ro(seqs + 1) : 0ull;
        rq2 = nseq > 2 ? ldg64_ro(seqs + 2) : 0ull;
        p_lit_end = p_out_end = 0;
        rep0 = cb->rep_init[0]; rep1 = cb->rep_init[1]; rep2 = cb->rep_init[2];
        const u32 lit_type = cb->lit_type;
        const u8* lit = lit_type >= 2 ? B.lit_pool + cb->lit_out : B.src + cb->src + cb->lit_src;
        lfill = lit_type == 1;
        fillw = lfill ? ldg8_ro(lit) * 0x01010101u : 0u;
        lal = (u32)((uintptr_t)lit & 15);
        lit16 = lit - lal;
        lp = 0; lbase = lw;
        bulkl = cb->bulk_lits != 0;
        tail = cb->lit_regen - cb->lit_used;
        have = false; is_tail = false; tail_done = false;
        st = ST_SEQ;
    };

    // ---- classifier: one entry of the current block.  Returns false when nothing could be queued.
    auto classify = [&]() -> bool {
        if (cnt >= D || st == ST_DONE) return false;
        if (st != ST_SEQ) { transition(); return true; }
        if (cnt && ex_epoch == cl_epoch && cc - c > LOOK_MAX) return false;
        if (!have) {
            if (ri < nseq) {
                const u64 rec = rq0;
                rq0 = rq1; rq1 = rq2;
                rq2 = ri + 3 < nseq ? ldg64_ro(seqs + ri + 3) : 0ull;
                ri++;
                const u32 lit_end = (u32)(rec >> 28) & LEN_MASK, out_end = (u32)(rec >> 46);
                rl = lit_end - p_lit_end;
                rmm = out_end - p_out_end - rl;
                p_lit_end = lit_end; p_out_end = out_end;
                off = (u32)rec & KZG_OFF_MASK;
                if (off & KZG_OFF_SYM) {
                    const u32 k = off & 3u, dd = (off & (KZG_OFF_SYM - 1u)) >> 2;
                    const u32 base = k == 0 ? rep0 : (k == 1 ? rep1 : rep2);
                    off = (base >= KZG_OFF_INVALID || base <= dd) ? KZG_OFF_INVALID : base - dd;
                }
                // (the sequence kernel guarantees non-decreasing lit_end / out_end within the block limits)
                if (off - 1u >= KZG_OFF_INVALID - 1u || off > cc + rl - cph) {      // no dictionary here: nothing before the frame
                    B.frames[cf].status = KZG_E_CORRUPTION;
                    B.frames[cf].fail_block = bk;
                    emit((K_END << 30), 0u, 0u, 0u);
                    st = ST_FRAME;
                    return true;
                }
                have = true;
            } else if (!tail_done && tail) {                // the literals after the last sequence, as a record without a match
                tail_done = true; is_tail = true;
                rl = tail; rmm = 0; off = 1; have = true;
            } else { bk++; st = ST_BLOCK; return true; }
        }
        if (rl >= KZG_BULK_MIN && (bulkl || is_tail)) {     // written by the bulk kernel
            hole(rl);
            lp += rl; rl = 0;
            lbase = lw - ((lal + lp) & ~15u);                 // the literal ring continues with the chunk that holds position lp
            if (rmm == 0) have = false;
            return true;
        }
        const u32 pll = mn(rl, KZG_X3_PIECE);
        const u32 P0 = lbase + lal + lp;                    // ring-space position of the first literal
        const u32 Pend = (P0 + pll + 15u) & ~15u;
        if (pll && Pend - (lit_done & ~15u) > KZG_X3_LITRING) return false;        // literal ring full
        u32 pml = 0, src = 0, spos = 0;
        bool far = false;
        const u32 mdst = cc + pll;
        if (rl <= KZG_X3_PIECE && rmm) {
            spos = mdst - off;
            if (off <= NEAR_MAX && (i32)(spos - Tv) >= 0) { pml = mn(rmm, KZG_X3_NEAR_PIECE); src = off; }
            else {
                pml = mn(mn(rmm, KZG_X3_PIECE), off);
                if ((i32)(spos - Tv) < 0) pml = mn(pml, Tv - spos);
                const u32 Feff = ex_epoch == cl_epoch ? F : cph;
                const bool avail = spos + pml <= Feff || (spos >= hole_lo && spos + pml <= hole_hi);
                if (!avail) {
                    if (cnt == 0 && c == e && !drain) want_flush = true;    // the executor is idle: a flush makes it so
                    return false;
                }
                // staging chunks: three more (plus what skipping the ring's end may waste) must be free
                if (((fw - fdone) & STAMP_MASK) > KZG_X3_FARCHUNKS - 6u) {
                    if (cnt == 0 && c == e) fdone = fw;                     // nothing queued or executing: all free
                    else return false;
                }
                far = true;
            }
        }
        const u32 fw0 = fw;
        // stage the literals (the chunks of the stream not requested yet; the ring's first 48 bytes are mirrored behind it)
#pragma unroll
        for (int q = 0; q < 3; q++) {
            if (pll && lw < Pend) {
                const u32 o = lw & (KZG_X3_LITRING - 1u);
                if (lfill) {
                    sts128(litL + o, make_uint4(fillw, fillw, fillw, fillw));
                    if (o < 48u) sts128(litL + KZG_X3_LITRING + o, make_uint4(fillw, fillw, fillw, fillw));
                } else {
                    const u8* g = lit16 + (lw - lbase);
                    cp_async16_if(true, litL + o, g);
                    cp_async16_if(o < 48u, litL + KZG_X3_LITRING + o, g);
                }
                lw += 16u;
            }
        }
        if (far) {
            const u8* g = cgal + spos;
            const u32 a = (u32)((uintptr_t)g & 15);
            const u32 nch = (a + pml + 15u) >> 4;           // <= 3
            u32 idx = fw % KZG_X3_FARCHUNKS;
            if (idx + nch > KZG_X3_FARCHUNKS) { fw += KZG_X3_FARCHUNKS - idx; idx = 0; }    // contiguous: skip the ring's end
            const u32 fa = farL + (idx << 4);
            g -= a;
            cp_async16_if(true, fa, g);
            cp_async16_if(nch > 1, fa + 16, g + 16);
            cp_async16_if(nch > 2, fa + 32, g + 32);
            fw += nch;
            src = fa + a;
        }
        emit((K_SEQ << 30) | (far ? E_FAR : 0u) | pll, pml, src | ((fw0 & STAMP_MASK) << 18), P0);
        lp += pll; rl -= pll; rmm -= pml;
        cc = mdst + pml;
        if (rl == 0 && rmm == 0) have = false;
        return true;
    };

    // ---- the warp's flush: head bytes up to a 16-byte boundary by their owner, then complete 16-byte chunks, four lanes
    // per frame; lanes that are draining (end of frame, a hole ahead) then write their ragged tail and move on.
    auto flush_event = [&]() {
        w_sync();
        const u32 up = (F + 15u) & ~15u;
        if ((F & 15u) && (c >= up || drain)) {
            const u32 hi = mn(up, c);
            for (u32 X = F; X < hi; X++) stg8(gal + X, lds8(ringL + (((X >> 2) & M) << 7) + (X & 3u)));
            F = hi;
        }
        const u32 glo = (u32)(uintptr_t)gal, ghi = (u32)((u64)(uintptr_t)gal >> 32);
#pragma unroll 1
        for (u32 r = 0; r < 4; r++) {
            const u32 T = (r << 3) + (lane >> 2), q = lane & 3u;
            const u32 FT = w_shfl(F, T), cT = w_shfl(c, T);
            const u32 gl = w_shfl(glo, T), gh = w_shfl(ghi, T);
            u8* galT = (u8*)(uintptr_t)(((u64)gh << 32) | gl);
            const u32 ringT = sb + L::RING + (T << 2);
            if (!(FT & 15u)) {
                for (u32 X = FT + (q << 4); X + 16u <= cT; X += 64u) {
                    const u32 a = ringT + (((X >> 2) & M) << 7);
                    stg128(galT + X, make_uint4(lds32(a), lds32(a + 128), lds32(a + 256), lds32(a + 384)));
                }
            }
        }
        if (!(F & 15u) && (c & ~15u) > F) F = c & ~15u;
        if (drain) {
            for (u32 X = F; X < c; X++) stg8(gal + X, lds8(ringL + (((X >> 2) & M) << 7) + (X & 3u)));
            if (drain == 2) c += drain_n;
            F = m = e = c;
            drain = 0;
        }
        fence_gpu();                    // far match sources are read back from L2 by other lanes
        w_sync();
        want_flush = false;
    };

    u32 idle_rounds = 0;
    for (;;) {
        // 1. classifier: at most one entry per lane
        bool progress = classify();

        // one cp.async group per round for the whole warp; the group of KZG_X3_AGE rounds ago has landed after this
        cp_commit();
        cp_wait<KZG_X3_AGE - 1>();

        // 2. executor: next entry, once it is old enough
        if (c == e && cnt && !drain) {
            const uint4 en = lds128(fifoL + ((head & (KZG_X3_FD - 1u)) << 9));
            if (((it - (en.y >> 18)) & STAMP_MASK) >= (u32)KZG_X3_AGE) {
                head++; cnt--;
                progress = true;
                const u32 kind = en.x >> 30;
                if (kind == K_SEQ) {
                    const u32 ll = en.x & LEN_MASK;
                    m = c + ll; e = m + (en.y & LEN_MASK);
                    lit0 = litL + (en.w & (KZG_X3_LITRING - 1u)) - c;
                    lit_done = en.w;                        // literal ring space below this entry's literals is free
                    fdone = en.z >> 18;                     // ... and so are the staging chunks allocated before it
                    isnear = !(en.x & E_FAR);
                    const u32 z = en.z & LEN_MASK;
                    dist = z; sfix = m - z;
                    msx = z - m;
                } else if (kind == K_START) {
                    gal = (u8*)(uintptr_t)(((u64)en.w << 32) | en.z);
                    c = m = e = F = en.x & LEN_MASK;
                    ex_epoch++;
                } else {
                    drain = kind == K_END ? 1u : 2u;
                    drain_n = en.x & LEN_MASK;
                    want_flush = true;
                }
            }
        }

        // 3. executor: one pass of up to 16 bytes
        if (c < e) {
            progress = true;
            const bool inlit = c < m;
            const u32 k = c & 3u;
            u32 n = mn((inlit ? m : e) - c, 16u - k);
            n = mn(n, W - (c & (W - 1u)));                  // the destination does not wrap inside a pass
            const bool ringsrc = !inlit && isnear;
            const u32 s = c - dist;
            if (ringsrc) { n = mn(n, dist); n = mn(n, W - (s & (W - 1u))); }
            const u32 A = (inlit ? lit0 : msx) + c;
            const u32 a0 = ringsrc ? ringL + (((s >> 2) & M) << 7) : (A & ~3u);
            const u32 r8 = ((ringsrc ? s : A) & 3u) << 3;
            const u32 stp = ringsrc ? 128u : 4u;
            const u32 w0 = lds32(a0), w1 = lds32(a0 + stp), w2 = lds32(a0 + 2u * stp), w3 = lds32(a0 + 3u * stp), w4 = lds32(a0 + 4u * stp);
            const u32 S0 = funnel_r(w0, w1, r8), S1 = funnel_r(w1, w2, r8), S2 = funnel_r(w2, w3, r8), S3 = funnel_r(w3, w4, r8);
            const u32 da = ringL + (((c >> 2) & M) << 7);
            const u32 k8 = k << 3;
            const u32 oldx = lds32(da) << ((32u - k8) & 31u);           // the word's first k bytes, moved to the top
            sts32(da, funnel_l(oldx, S0, k8));
            sts32(da + 128, funnel_l(S0, S1, k8));
            sts32(da + 256, funnel_l(S1, S2, k8));
            sts32(da + 384, funnel_l(S2, S3, k8));
            c += n;
            if (ringsrc && dist < 16u && c - sfix >= 2u * dist) dist <<= 1;    // the period has unrolled once more
        }

        const bool alive = st != ST_DONE || cnt || c < e || drain;
        const u32 fl = w_or((alive ? 1u : 0u) | ((want_flush || c - F >= KZG_X3_FLUSH) ? 2u : 0u) | (progress ? 4u : 0u));
        if (fl & 2u) flush_event();
        if (!(fl & 1u)) break;
        it++;
        if (fl & 6u) idle_rounds = 0;
        else if (++idle_rounds > 100000u) {     // cannot happen; never spin forever on a GPU
            if (alive && st != ST_DONE) B.frames[cf].status = KZG_E_GENERIC;
            break;
        }
    }
    cp_wait<0>();
}

#if !defined(KZG_SIMT_EMU)
template <u32 W>
__global__ void __launch_bounds__(32) kzg_exec3_kernel(KzgBatch B) {
    extern __shared__ uint4 smem4[];
    if (B.ctr->overflow & 3u) return;
    kzg_exec3_warp<W>(B, (u32)__cvta_generic_to_shared(smem4), threadIdx.x & 31u);
}
#endif
