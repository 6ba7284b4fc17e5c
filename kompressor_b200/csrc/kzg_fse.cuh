// Sequence stage (s4) for sm_100a: FSE decode-table build + the ZSTD_decodeSequence state machine,
// one GPU thread per compressed block, 32 blocks per warp.  Device-only code.
//
// The state machine is a serial dependency chain per block (state -> table cell -> bit count -> bits
// -> next state).  A single warp can only issue that chain one dependent instruction after another, so
// the kernel is bound by instructions per sequence, not by bandwidth.  The design therefore minimises
// and splits the per-sequence instruction stream:
//   * warp specialisation: every group of 32 blocks is served by a PAIR of warps on two different
//     schedulers.  The "chain" warp does only what the next state depends on (table lookups, bit
//     counts, state update, bit position); the "post" warp turns the raw fields into values, keeps the
//     repeat-offset history, accumulates lit_end / out_end, packs and stores the 8-byte records.  The
//     two meet through a double-buffered shared-memory queue (4 sequences per hand-over) guarded by
//     named barriers (bar.arrive / bar.sync), never by spinning;
//   * tables live in shared memory, interleaved by lane: cell i of lane L is the u16 at byte
//     (i << 6) + (L << 1) of the pair's table area, so a lookup costs one multiply-add and lanes collide
//     at most pairwise (two lanes share a 32-bit bank column); the tables are built once per block by
//     kzg_fsetab_kernel (warp-parallel) and only copied in here, LL and OF by the chain warp, ML by the post warp;
//   * the backward bitstream is staged by cp.async (LDGSTS, no destination register, so no per-warp
//     scoreboard stall) into a per-lane ring of eight 16-byte chunks (128 contiguous bytes per lane);
//     copies are issued only at warp-uniform points (every 4 sequences), sized so that the ring always
//     covers the most bits four sequences can consume (4 x 89);
//   * every sequence re-reads a 64-bit window at its current bit position from that ring (3 LDS + 2
//     funnel shifts, off the critical chain) and all six bit fields come out of it with clamped funnel
//     shifts: no refill branches, no window register rotation; the loop body is straight-line code and
//     lanes that ran out of sequences keep computing on frozen positions;
//   * sequences carrying more than 32 extra bits at once (possible only with huge offsets and lengths)
//     make the lane fall back to a general decoder for that block after the loop.
// What libzstd does here is ZSTD_decodeSequence / ZSTD_buildFSETable, reached from the reference through
// ZSTD_decompressStream (kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:178).
#pragma once
#include "kzg_stages.cuh"
#include "kzg_huf.cuh"

#define KZG_FSE_PAIRS 2                                     // warp pairs per CTA (one CTA per SM)
#define KZG_FSE_WARPS (2 * KZG_FSE_PAIRS)
#define KZG_FSE_CELLS 1280                                  // LL 512 | ML 512 | OF 256
#define KZG_FSE_TAB_BYTES (KZG_FSE_CELLS * 64)              // per pair: 32 lanes x 1280 cells x 2 B
#define KZG_FSE_RING_CHUNKS 8
#define KZG_FSE_RING_BYTES (KZG_FSE_RING_CHUNKS * 512)      // per pair: 32 lanes x (8 chunks x 16 B, contiguous per lane)
#define KZG_FSE_SERVICE 4                                   // sequences between ring service points
#ifndef KZG_FSE_HAND
#define KZG_FSE_HAND 8                                      // sequences per hand-over (a multiple of KZG_FSE_SERVICE)
#endif
#ifndef KZG_FSE_INFLIGHT
#define KZG_FSE_INFLIGHT 3                                  // copy groups (one per ring service) that may still be in flight at a service point: 1..3
#endif
#define KZG_FSE_Q_BYTES (2 * KZG_FSE_HAND * 512)            // per pair: 2 buffers x 8 sequences x 32 lanes x 16 B
#define KZG_FSE_MISC_BYTES 512                              // per pair: ticket, trip count, per-lane hand-over words
#define KZG_FSE_PAIR_SMEM (KZG_FSE_TAB_BYTES + KZG_FSE_RING_BYTES + KZG_FSE_Q_BYTES + KZG_FSE_MISC_BYTES)
#define KZG_FSE_CODE_WORDS 192                              // 36 LL + 53 ML packed code entries, padded to 128 words; then (from byte 512,
                                                            // aligned to 64) the extra-bit counts alone, one BYTE per code: 64 B LL, 64 B ML
#define KZG_FSE_ALIGN_SLACK 128                             // the kernel aligns its area to 128 bytes itself
#define KZG_FSE_SMEM (KZG_FSE_CODE_WORDS * 4 + KZG_FSE_PAIRS * KZG_FSE_PAIR_SMEM)
// Up to three more warps ride along in every CTA and decode Huffman literals (kzg_huf_warp_loop) in the shared memory
// and the issue slots the four sequence warps leave free -- where they can finish all of them meanwhile (see the kernel).
// Literal warps are the odd warps 5, 7, ... (the even ones in between exit at once): they share a scheduler with a post
// warp, not with a chain warp.  With the two-level Huffman tables a literal warp needs 13 KiB, so up to three fit.
#define KZG_FSE_LIT_WARPS_MAX 3
#define KZG_FSE_THREADS_OF(nlit) (KZG_FSE_WARPS * 32 + 64 * (nlit))
#define KZG_FSE_THREADS KZG_FSE_THREADS_OF(KZG_FSE_LIT_WARPS_MAX)
#define KZG_FSE_SMEM_TOTAL_OF(nlit) (KZG_FSE_ALIGN_SLACK + KZG_FSE_SMEM + (nlit) * KZG_HUF_WARP_SMEM)
#define KZG_FSE_SMEM_TOTAL KZG_FSE_SMEM_TOTAL_OF(KZG_FSE_LIT_WARPS_MAX)
#define KZG_FSE_DONE_WORD 127                               // in the code-word area: pairs of this CTA that have finished
#ifndef KZG_FSE_SEQ_CYCLES
#define KZG_FSE_SEQ_CYCLES 200ull                           // one sequence of the chain, in SM clocks (measured: 175 alone, 190 with riders; 170 .. 220 give the same times on log text: the literal warps are done before any of these deadlines)
#endif
#ifndef KZG_FSE_WIDE_LIMIT
#define KZG_FSE_WIDE_LIMIT 32u                              // most extra bits per sequence the fast loop extracts (tests lower it)
#endif

namespace kzg_fse {

__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u32 lds16(u32 a) { u16 v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts16(u32 a, u32 v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((u16)v)); }
__device__ __forceinline__ u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ u32 lds8(u32 a) { u32 v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(u32 a, u32 x, u32 y, u32 z, u32 w) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void lds128(u32 a, u32& x, u32& y, u32& z, u32& w) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a) : "memory");
}
// Named barriers over the 64 threads of a warp pair: sync waits, arrive only signals.
__device__ __forceinline__ void bar_sync(u32 id) { asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive(u32 id) { asm volatile("bar.arrive %0, 64;" :: "r"(id) : "memory"); }
__device__ __forceinline__ void cp_async16(u32 dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
// Predicated form: one instruction, no branch around it.
__device__ __forceinline__ void cp_async16_if(bool p, u32 dst, const void* src) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" :: "r"(dst), "l"(src), "r"((u32)p) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Copies one built table (2^log cells, contiguous in the workspace) into the lane's column of the interleaved
// shared-memory layout: 16-byte loads, eight 2-byte scattered stores each.  Iterations are independent.
__device__ __forceinline__ void load_table(u32 tab, const u16* g, u32 log) {
    const u32 nvec = max(1u << log, 8u) >> 3;
    const uint4* g4 = reinterpret_cast<const uint4*>(g);
#pragma unroll 4
    for (u32 q = 0; q < nvec; q++) {
        const uint4 v = __ldg(g4 + q);
        const u32 a = tab + (q << 9);
        sts16(a, v.x); sts16(a + 64, v.x >> 16); sts16(a + 128, v.y); sts16(a + 192, v.y >> 16);
        sts16(a + 256, v.z); sts16(a + 320, v.z >> 16); sts16(a + 384, v.w); sts16(a + 448, v.w >> 16);
    }
}

// Generic (slow, rare) backward bit read straight from global memory: n <= 32 bits below position P.
__device__ __forceinline__ u32 slow_bits(const u32* base32, i32 P, u32 n) {
    if (n == 0) return 0;
    const i32 k = P >> 5;
    const u32 q = (u32)P & 31u;
    const u32 w1 = k >= 0 ? base32[k] : 0u, w0 = k >= 1 ? base32[k - 1] : 0u;
    return __funnelshift_r(w0, w1, q) >> (32 - n);
}

struct Result { u32 status, regen, lit_used, rep0, rep1, rep2, bulk_lits; };

// Per-block decode parameters shared by the fast loop and the general fallback.
struct Job {
    const u32* base32;      // 16-byte aligned chunk holding the first byte of the bitstream
    u64* out;               // sequence records
    i32 P0, lowbits;        // bit position above the first unread bit / of the stream's first bit
    u32 nseq, lit_regen, block_max;
    u32 logLL, logOF, logML;
    u32 rep0, rep1, rep2;   // encoded repeat offsets at block start
};

__device__ __forceinline__ void finish_result(Result& res, const Job& j, i32 P, u32 lit_end, u32 out_end, u32 ovf,
                                              u32 rep0, u32 rep1, u32 rep2) {
    // P only decreases, lit_end / out_end only grow: checking the final values covers every step.
    if (P == j.lowbits && (ovf >> 18) == 0 && lit_end <= j.lit_regen && out_end <= j.block_max) {
        const u32 regen = out_end + (j.lit_regen - lit_end);
        if (regen <= j.block_max) {
            res.status = KZG_OK; res.regen = regen; res.lit_used = lit_end;
            res.rep0 = rep0; res.rep1 = rep1; res.rep2 = rep2;
        }
    }
}

// General decoder for one block: any field widths (up to 89 bits per sequence), bits read from global
// memory.  Only used for blocks whose sequences carry more than 32 extra bits at once, which the fast
// loop does not handle; one lane runs it after the fast loop, so speed does not matter.
__device__ __noinline__ void slow_decode(const Job& j, u32 tLL, u32 tML, u32 tOF, u32 s_llx, u32 s_mlx, Result& res) {
    i32 P = j.P0;
    u32 sLL = slow_bits(j.base32, P, j.logLL); P -= (i32)j.logLL;
    u32 sOF = slow_bits(j.base32, P, j.logOF); P -= (i32)j.logOF;
    u32 sML = slow_bits(j.base32, P, j.logML); P -= (i32)j.logML;
    if (P < j.lowbits) return;
    u32 rep0 = j.rep0, rep1 = j.rep1, rep2 = j.rep2, lit_end = 0, out_end = 0, ovf = 0;
    res.bulk_lits = 0;
    for (u32 i = 0; i < j.nseq; i++) {
        const u32 eLL = lds16(tLL + (sLL << 6)), eML = lds16(tML + (sML << 6)), eOF = lds16(tOF + (sOF << 6));
        const u32 ofc = eOF & 63u;
        const u32 pLL = lds32(s_llx + ((eLL & 63u) << 2)), pML = lds32(s_mlx + ((eML & 63u) << 2));
        const u32 ofv = (1u << ofc) + slow_bits(j.base32, P, ofc); P -= (i32)ofc;
        const u32 ml = (pML & 0xFFFFFFu) + slow_bits(j.base32, P, pML >> 24); P -= (i32)(pML >> 24);
        const u32 ll = (pLL & 0xFFFFFFu) + slow_bits(j.base32, P, pLL >> 24); P -= (i32)(pLL >> 24);
        u32 off;
        if (ofv > 3) {
            off = min(ofv - 3u, (u32)KZG_OFF_INVALID);
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
        ovf |= out_end;
        if (ll >= KZG_BULK_MIN) res.bulk_lits += ll;
        j.out[i] = kzg_pack_seq(off, lit_end & 0x3FFFFu, out_end & 0x3FFFFu);
        if (i + 1 < j.nseq) {
            const u32 nLL = eLL >> 6, nML = eML >> 6, nOF = eOF >> 6;
            const u32 bLL = j.logLL - (u32)kzg_highbit(nLL), bML = j.logML - (u32)kzg_highbit(nML), bOF = j.logOF - (u32)kzg_highbit(nOF);
            sLL = (nLL << bLL) - (1u << j.logLL) + slow_bits(j.base32, P, bLL); P -= (i32)bLL;
            sML = (nML << bML) - (1u << j.logML) + slow_bits(j.base32, P, bML); P -= (i32)bML;
            sOF = (nOF << bOF) - (1u << j.logOF) + slow_bits(j.base32, P, bOF); P -= (i32)bOF;
        }
        if (P < j.lowbits) return;
    }
    finish_result(res, j, P, lit_end, out_end, ovf, rep0, rep1, rep2);
}

}  // namespace kzg_fse

// Queue entry of one sequence of one lane (16 bytes): {hi, lo, eLL | eML << 16, ofc}.
//   hi:lo  64-bit window at the sequence's bit position        eLL / eML  the LL / ML table cells (code in bits 0-5)
__global__ void __launch_bounds__(KZG_FSE_THREADS, 1) kzg_fse_kernel(KzgBatch B) {
    using namespace kzg_fse;
    extern __shared__ uint4 smem4_raw[];
    // the area starts on a 128-byte boundary (KZG_FSE_ALIGN_SLACK bytes are allocated for that): a lane's 128-byte ring is
    // then aligned, and a ring address is one LOP3 -- (offset & 0x7C) | ring -- instead of an AND and an add
    uint4* const smem4 = reinterpret_cast<uint4*>(reinterpret_cast<u8*>(smem4_raw) + ((0u - smem_addr(smem4_raw)) & 127u));
    u32* llx = reinterpret_cast<u32*>(smem4);
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 pair = warp >> 1;
    const bool chain = (warp & 1u) == 0;                   // even warp: chain, odd warp: post
    for (u32 k = threadIdx.x; k < 128u; k += blockDim.x) llx[k] = k < 89 ? B.code_tables[k] : 0u;
    // the extra-bit counts of the LL / ML codes as bytes, each table 64-byte aligned: its address is one LOP3 of the cell
    const u32 s_llb = (smem_addr(llx) + 512u + 63u) & ~63u, s_mlb = s_llb + 64u;
    if (threadIdx.x < 64u) {
        const u32 k = threadIdx.x;
        asm volatile("st.shared.u8 [%0], %1;" :: "r"(s_llb + k), "r"(k < 36u ? B.code_tables[k] >> 24 : 0u) : "memory");
        asm volatile("st.shared.u8 [%0], %1;" :: "r"(s_mlb + k), "r"(k < 53u ? B.code_tables[36u + k] >> 24 : 0u) : "memory");
    }
    __syncthreads();
    if (warp >= KZG_FSE_WARPS && !(warp & 1u)) return;
    const long long t_start = clock64();
    // Whether literal warps ride in this launch, and until when they take new tickets (0: they do not ride); see below.
    auto ride_deadline = [&]() -> long long {
        const u64 nsb = B.ctr->n_sblocks, nhb = B.ctr->n_hblocks, nsq = B.ctr->seq_used;
        if (nsb == 0 || nhb == 0 || nsq < nsb * 1024u || B.lit_ride == 2u) return 0;
        const u64 lanes = (u64)gridDim.x * (KZG_FSE_PAIRS * 32u);
        const u64 est_seq = ((nsb + lanes - 1) / lanes) * (nsq / nsb);
        const u64 riders = (u64)gridDim.x * ((blockDim.x - KZG_FSE_WARPS * 32u) / 64u);
        const u64 est_tkt = (B.ctr->lit_used / nhb / 4u) * 3u / 4u;                 // one ticket: the streams of 8 blocks side by side
        // tickets the riders can finish before the sequence stage ends, against all of them: at least four fifths, or none
        // (measured: riders that take 87 % leave kzg_huf_kernel a shorter, emptier launch and win 2-5 %; 45 % loses 2 %)
        const u64 fit = est_tkt ? est_seq / est_tkt : 0xFFFFull;
        if (B.lit_ride == 0u && fit * riders * 5u < ((nhb + 7) / 8) * 4u) return 0;
        return t_start + (long long)((est_seq > est_tkt ? est_seq - est_tkt : 0ull) * KZG_FSE_SEQ_CYCLES) + 1;
    };
    if (warp > KZG_FSE_WARPS) {                             // a literal warp
        // Both stages are serial chains, so the riders are all or nothing: whatever they leave takes kzg_huf_kernel one
        // full chain however little it is, and riders cost the sequence warps ~7 % (shared-memory traffic).  Measured on
        // 1 GiB of 64 KiB frames: JSON 0.995 + 0.417 ms with riders that cannot finish, 0.740 + 0.430 without; log text,
        // whose literals are half as long, finishes under the sequence stage: 0.854 + 0.012 against 0.793 + 0.224.  So
        // they ride only where they can take (nearly) all the literal tickets before the sequence stage ends, by an estimate
        // from the batch's counters in units of one sequence of the chain (KZG_FSE_SEQ_CYCLES; a rider's literal symbol
        // takes ~3/4 of that), and no rider takes a new ticket it could not finish by then (clock64 deadline; every SM
        // starts its one CTA at the same time).  Riders on SMs whose pairs have no block ride too.
        const long long deadline = ride_deadline();
        if (!deadline) return;
        kzg_huf_warp_loop<KZG_HUF_L1_BITS, KZG_HUF_L2CAP, false>(B, reinterpret_cast<u8*>(smem4) + KZG_FSE_SMEM + ((warp - KZG_FSE_WARPS - 1u) >> 1) * KZG_HUF_WARP_SMEM, lane, deadline, 0xFFFFFFFFu);
        return;
    }
    const u32 s_llx = smem_addr(llx), s_mlx = s_llx + 36 * 4;
    const u32 s_pair = s_llx + KZG_FSE_CODE_WORDS * 4 + pair * KZG_FSE_PAIR_SMEM;
    const u32 s_tab = s_pair + (lane << 1);
    const u32 s_ring = s_pair + KZG_FSE_TAB_BYTES + (lane << 7);
    // queue slot of one sequence: 16 bytes per lane, the lanes back to back (a warp's access covers 512 contiguous bytes:
    // no bank conflicts; entries of 32 bytes per lane made every access 2-way / 8-way conflicted)
    const u32 s_q = s_pair + KZG_FSE_TAB_BYTES + KZG_FSE_RING_BYTES + (lane << 4);
    const u32 s_misc = s_pair + KZG_FSE_TAB_BYTES + KZG_FSE_RING_BYTES + KZG_FSE_Q_BYTES;
    const u32 s_lane = s_misc + 64 + (lane << 2);           // per-lane hand-over word
    const u32 tLL = s_tab, tML = s_tab + (512u << 6), tOF = s_tab + (1024u << 6);
    // barrier ids of this pair (0 is __syncthreads)
    const u32 BAR_PAIR = 1 + pair * 5, BAR_FULL = 2 + pair * 5, BAR_EMPTY = 4 + pair * 5;
    if (B.ctr->overflow & 3u) return;
    const u32 n = B.ctr->n_sblocks;

    for (;;) {
        if (chain && lane == 0) sts32(s_misc, atomicAdd(&B.ctr->ticket_fse, 1u));
        bar_sync(BAR_PAIR);
        const u32 g = lds32(s_misc);
        if (g * 32 >= n) { if (chain && lane == 0) atomicAdd(llx + KZG_FSE_DONE_WORD, 1u); break; }
        const u32 si = g * 32 + lane;
        KzgCBlock* cb = nullptr;
        Job j;
        j.base32 = nullptr; j.out = nullptr; j.P0 = 0; j.lowbits = 0; j.nseq = 0; j.lit_regen = 0; j.block_max = 0;
        j.logLL = j.logOF = j.logML = 0;
        j.rep0 = 1; j.rep1 = 4; j.rep2 = 8;
        if (si < n) {
            cb = &B.cblocks[B.slist[si]];
            if (cb->status_seq == KZG_OK && cb->status_lit == KZG_OK) j.nseq = cb->n_seq;
        }
        // ---- tables (built by kzg_fsetab_kernel) into the lane's column: the chain warp loads LL and OF, the
        // post warp ML.  Idle lanes get one-cell tables so the straight-line loop stays in bounds.
        if (j.nseq) {
            const u32 o0 = cb->fse_owner[0], o1 = cb->fse_owner[1], o2 = cb->fse_owner[2];
            const KzgFseDesc* d0 = kzg_fse_desc_of(B, o0);
            const KzgFseDesc* d1 = kzg_fse_desc_of(B, o1);
            const KzgFseDesc* d2 = kzg_fse_desc_of(B, o2);
            j.logLL = d0->kind[0] == 1 ? 0u : (u32)d0->log[0];
            j.logOF = d1->kind[1] == 1 ? 0u : (u32)d1->log[1];
            j.logML = d2->kind[2] == 1 ? 0u : (u32)d2->log[2];
            const u16* gt = B.fse_tab + (size_t)si * KZG_FSE_CELLS;
            if (chain) { load_table(tLL, gt, j.logLL); load_table(tOF, gt + 1024, j.logOF); }
            else load_table(tML, gt + 512, j.logML);
            j.lit_regen = cb->lit_regen;
            j.block_max = B.frames[cb->frame].block_max;
            j.out = B.seq_pool + cb->seq_out;
            if (!cb->first_in_frame) { j.rep0 = KZG_OFF_SYM | 0; j.rep1 = KZG_OFF_SYM | 1; j.rep2 = KZG_OFF_SYM | 2; }
            else if (B.dict_entropy) { j.rep0 = B.dict_rep[0]; j.rep1 = B.dict_rep[1]; j.rep2 = B.dict_rep[2]; }
        } else if (chain) { sts16(tLL, 1u << 6); sts16(tOF, 1u << 6); sts16(tML, 1u << 6); }
        Result res;
        res.status = KZG_E_CORRUPTION; res.regen = 0; res.lit_used = 0; res.bulk_lits = 0;
        res.rep0 = j.rep0; res.rep1 = j.rep1; res.rep2 = j.rep2;

        // ---- bitstream set-up (chain warp); the effective sequence count goes to the post warp
        const u8* base = nullptr;
        i32 next_c = -1, ready_lo = 0, pend1 = 0, pend2 = 0, pend3 = 0, pend4 = 0, P = 0;
        u32 sLL = 0, sOF = 0, sML = 0;
        u32 nseq = j.nseq;
        bar_sync(BAR_PAIR);                                  // tables complete (all three, both warps)
        if (chain) {
            if (nseq) {
                const u8* start = B.src + cb->src + cb->seq_src;
                const u32 size = cb->seq_size;
                const u32 last = size ? start[size - 1] : 0u;
                if (last == 0) nseq = 0;                   // no padding marker: corrupt (status stays CORRUPTION)
                else {
                    base = (const u8*)((uintptr_t)start & ~(uintptr_t)15);
                    j.base32 = reinterpret_cast<const u32*>(base);
                    j.lowbits = (i32)(8u * (u32)(start - base));
                    j.P0 = j.lowbits + (i32)(8u * (size - 1)) + kzg_highbit(last);
                    next_c = j.P0 >> 7;
                    for (int r = 0; r < KZG_FSE_RING_CHUNKS; r++)
                        if (next_c >= 0) { cp_async16(s_ring + (((u32)next_c & 7u) << 4), base + ((size_t)next_c << 4)); next_c--; }
                }
            }
            cp_commit();
            cp_wait<0>();
            ready_lo = pend1 = pend2 = pend3 = pend4 = next_c + 1;
            // Lanes without a block keep executing the loop on a frozen position: give each its own ring word, or all of
            // them read the same bank of their (128-byte strided) rings three times per sequence -- a 32-way conflict
            // that doubles the time of a group with few blocks (small batches).
            P = nseq ? j.P0 : (i32)(lane << 5);
        }

        // 64-bit window [P-64, P) -> hi:lo from the lane's ring (word w of the stream sits at (w & 31) * 4)
#define KZG_FSE_WINDOW(hi, lo)                                                     \
        {                                                                          \
            const u32 t_ = (u32)(P >> 3), q_ = (u32)P & 31u;                       \
            const u32 w2_ = lds32(s_ring | (t_ & 0x7Cu));                          \
            const u32 w1_ = lds32(s_ring | ((t_ - 4u) & 0x7Cu));                   \
            const u32 w0_ = lds32(s_ring | ((t_ - 8u) & 0x7Cu));                   \
            hi = __funnelshift_r(w1_, w2_, q_);                                    \
            lo = __funnelshift_r(w0_, w1_, q_);                                    \
        }

        if (chain) {
            // initial states: LL, OF, ML (Appendix A.6)
            if (nseq) {
                u32 hi, lo;
                KZG_FSE_WINDOW(hi, lo);
                if (lo == 0xFFFFFFFFu && hi == 0u) hi = 0u;    // (lo is unused here)
                sLL = __funnelshift_lc(hi, 0u, j.logLL);
                const u32 x1 = hi << j.logLL;
                sOF = __funnelshift_lc(x1, 0u, j.logOF);
                const u32 x2 = x1 << j.logOF;
                sML = __funnelshift_lc(x2, 0u, j.logML);
                P -= (i32)(j.logLL + j.logOF + j.logML);
                if (P < j.lowbits) nseq = 0;
            }
            sts32(s_lane, nseq);
            u32 nmax = nseq;
#pragma unroll
            for (int o = 16; o; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, o));
            if (lane == 0) sts32(s_misc + 4, nmax);
        }
        bar_sync(BAR_PAIR);
        const u32 nmax = lds32(s_misc + 4);
        nseq = lds32(s_lane);
        const u32 nint = (nmax + KZG_FSE_HAND - 1) / KZG_FSE_HAND;     // hand-overs in this round

        if (chain) {
            // =================================================================== chain warp
            const u32 szLL = 1u << j.logLL, szOF = 1u << j.logOF, szML = 1u << j.logML;
            const u32 kLL = j.logLL - 25u, kOF = j.logOF - 25u, kML = j.logML - 25u;   // nbBits = k + clz(cell): a cell is next << 6 | symbol, next >= 1
            bool wide = false;                   // a sequence with more than 32 extra bits was seen
            // The loop is one dependency chain per sequence (cell -> extra-bit counts -> E -> state bits -> next cell),
            // and every dependent instruction on it costs a lone warp ~12 clocks, a shared-memory load ~32.  So the chain
            // carries the ADDRESS of each state's cell, not the state (the part of the next address that does not depend on
            // the bits read -- table base + ((next << nb) - size) * 64 -- is computed beside the chain), the extra-bit
            // counts come from byte tables whose address is a single LOP3 of the cell, and E is a single three-input add:
            // two loads and five instructions between one LL cell and the next (it was two loads and nine).
            u32 aLL = tLL + (sLL << 6), aML = tML + (sML << 6), aOF = tOF + (sOF << 6);
            const u32 zLL = tLL - (szLL << 6), zML = tML - (szML << 6), zOF = tOF - (szOF << 6);
            for (u32 t = 0; t < nint; t++) {
                const u32 buf = t & 1u;
#ifndef DBG_NOQ
                if (t >= 2) bar_sync(BAR_EMPTY + buf);            // the post warp is done with this buffer
#endif
#pragma unroll
                for (u32 h = 0; h < KZG_FSE_HAND / KZG_FSE_SERVICE; h++) {
                {
                    // ring service (warp-uniform point). Chunks issued two service points ago are complete now;
                    // up to 3 new chunks (384 bits >= 4 x 89, the most four sequences can consume) are requested,
                    // as far down as the ring allows (the slot of chunk c + 8 must be dead).
#ifdef DBG_NORING
                    if (t > 0x7fffffffu)
#endif
                    cp_wait<KZG_FSE_INFLIGHT>();
                    ready_lo = KZG_FSE_INFLIGHT == 1 ? pend2 : (KZG_FSE_INFLIGHT == 2 ? pend3 : pend4);
                    const i32 cP = P >> 7;
#ifdef DBG_NORING
                    if (t > 0x7fffffffu) {
#endif
                    const i32 want = min(min(next_c - (cP - KZG_FSE_RING_CHUNKS), next_c + 1), 3);
                    const u8* gsrc = base + ((i64)next_c << 4);
                    cp_async16_if(want > 0, s_ring + (((u32)next_c & 7u) << 4), gsrc);
                    cp_async16_if(want > 1, s_ring + (((u32)(next_c - 1) & 7u) << 4), gsrc - 16);
                    cp_async16_if(want > 2, s_ring + (((u32)(next_c - 2) & 7u) << 4), gsrc - 32);
                    next_c -= max(want, 0);
                    cp_commit();
                    pend4 = pend3;
                    pend3 = pend2;
                    pend2 = pend1;
                    pend1 = next_c + 1;
                    // lowest chunk the next four sequences can touch: 4 x 89 bits + 64-bit window + 31
                    const i32 need_lo = max((P - 451) >> 7, 0);
                    if (need_lo < ready_lo) { cp_wait<0>(); ready_lo = pend4 = pend3 = pend2 = pend1; }     // rare: wait for everything
#ifdef DBG_NORING
                    }
#endif
                }
                const u32 i0 = t * KZG_FSE_HAND + h * KZG_FSE_SERVICE;
                const u32 qb = s_q + buf * (KZG_FSE_HAND * 512) + h * (KZG_FSE_SERVICE * 512);
#pragma unroll
                for (u32 u = 0; u < KZG_FSE_SERVICE; u++) {
                    const u32 i = i0 + u;
                    const bool act = i < nseq, notlast = i + 1 < nseq;
                    const u32 eLL = lds16(aLL), eML = lds16(aML), eOF = lds16(aOF);
                    u32 hi, lo;
#ifdef DBG_NOWIN
                    hi = eLL * 2654435761u; lo = eML * 40503u;
#else
                    KZG_FSE_WINDOW(hi, lo);
#endif
                    const u32 ofc = eOF & 63u;
#ifdef DBG_NOCODE
                    const u32 llb = 0, mlb = 0;
#else
                    const u32 llb = lds8(s_llb | (eLL & 63u)), mlb = lds8(s_mlb | (eML & 63u));
#endif
                    // nbBits = log - floor(log2(next)) = log - 25 + clz(cell); next << 6 = the cell without its symbol
                    const u32 mLL = eLL & 0xFFC0u, mML = eML & 0xFFC0u, mOF = eOF & 0xFFC0u;
                    const u32 bLL = (u32)__clz((int)eLL) + kLL, bML = (u32)__clz((int)eML) + kML, bOF = (u32)__clz((int)eOF) + kOF;
                    const u32 E = ofc + mlb + llb;
                    wide |= act && E > KZG_FSE_WIDE_LIMIT;
#ifndef DBG_NOQ
                    sts128(qb + u * 512, hi, lo, eLL | (eML << 16), ofc);
#endif
                    // state bits sit right below the E extra bits (valid when E <= 32: E + 26 <= 58 bits)
                    const u32 X = __funnelshift_lc(lo, hi, E);
                    const u32 X2 = X << bLL, X3 = X2 << bML;
                    aLL = zLL + (mLL << bLL) + (__funnelshift_lc(X, 0u, bLL) << 6);
                    aML = zML + (mML << bML) + (__funnelshift_lc(X2, 0u, bML) << 6);
                    aOF = zOF + (mOF << bOF) + (__funnelshift_lc(X3, 0u, bOF) << 6);
                    P -= (i32)(act ? E + (notlast ? bLL + bML + bOF : 0u) : 0u);
                }
                }
#ifndef DBG_NOQ
                __threadfence_block();
                bar_arrive(BAR_FULL + buf);
#endif
            }
            cp_wait<0>();
            // hand the verdict on the bitstream to the post warp: bit 0 = consumed exactly, bit 1 = general path needed
            sts32(s_lane, (P == j.lowbits ? 1u : 0u) | (wide ? 2u : 0u));
            bar_sync(BAR_PAIR);
            if (cb && nseq && wide) {
                j.nseq = nseq;
                slow_decode(j, tLL, tML, tOF, s_llx, s_mlx, res);
                cb->status_seq = res.status;
                cb->regen = res.regen;
                cb->lit_used = res.lit_used;
                cb->rep_final[0] = res.rep0; cb->rep_final[1] = res.rep1; cb->rep_final[2] = res.rep2;
                cb->bulk_lits = res.bulk_lits >= KZG_BULK_SUM ? 1u : 0u;
            }
        } else {
            // =================================================================== post warp
            u32 rep0 = j.rep0, rep1 = j.rep1, rep2 = j.rep2;
            u32 lit_end = 0, out_end = 0, ovf = 0, longsum = 0;
            u64* out = j.out;
            for (u32 t = 0; t < nint; t++) {
                const u32 buf = t & 1u;
#ifndef DBG_NOQ
                bar_sync(BAR_FULL + buf);
#endif
                const u32 i0 = t * KZG_FSE_HAND;
                const u32 qb = s_q + buf * (KZG_FSE_HAND * 512);
#pragma unroll
                for (u32 u = 0; u < KZG_FSE_HAND; u++) {
                    const u32 i = i0 + u;
                    const bool act = i < nseq;
                    u32 hi, lo, cells, ofc;
                    lds128(qb + u * 512, hi, lo, cells, ofc);
                    const u32 pLL = lds32(s_llx + ((cells & 63u) << 2)), pML = lds32(s_mlx + ((cells >> 14) & 252u));
                    const u32 llb = pLL >> 24, mlb = pML >> 24;
                    const u32 ofx = __funnelshift_lc(hi, 0u, ofc);
                    const u32 mlx_ = __funnelshift_lc(__funnelshift_l(lo, hi, ofc), 0u, mlb);
                    const u32 llx_ = __funnelshift_lc(__funnelshift_lc(lo, hi, ofc + mlb), 0u, llb);
                    const u32 ofv = (1u << ofc) + ofx;
                    const u32 ml = act ? (pML & 0xFFFFFFu) + mlx_ : 0u;
                    const u32 ll = act ? (pLL & 0xFFFFFFu) + llx_ : 0u;
                    // repeat-offset history (Appendix A.6) with selects.  "rep0 - 1": a symbolic value counts its
                    // decrement in bits [2,27) (+4), a concrete one just loses 1 (0 is rejected at execution).
                    const bool isrep = ofc < 2u;                         // ofv in {1, 2, 3}
                    const u32 idx = ofv - 1u + (ll == 0u ? 1u : 0u);     // meaningful when isrep
                    const u32 dec = rep0 - 1u + (rep0 >> 27) * 5u;
                    u32 cand = idx == 1u ? rep1 : rep0;
                    cand = idx == 2u ? rep2 : cand;
                    cand = idx == 3u ? dec : cand;
                    const u32 off = isrep ? cand : min(ofv - 3u, (u32)KZG_OFF_INVALID);
                    rep2 = ((isrep && idx <= 1u) || !act) ? rep2 : rep1;
                    rep1 = ((isrep && idx == 0u) || !act) ? rep1 : rep0;
                    rep0 = act ? off : rep0;
                    lit_end += ll;
                    out_end += ll + ml;
                    longsum += ll >= KZG_BULK_MIN ? ll : 0u;
                    ovf |= out_end;           // one sequence adds < 2^18, so any overflow past 2^18 leaves a high bit here
                    if (act) out[i] = kzg_pack_seq(off, lit_end & 0x3FFFFu, out_end & 0x3FFFFu);
                }
                if (t + 2 < nint) {          // matched by the chain warp's wait two hand-overs later
#ifndef DBG_NOQ
                    __threadfence_block();   // the queue reads above are performed before the buffer is handed back
                    bar_arrive(BAR_EMPTY + buf);
#endif
                }
            }
            bar_sync(BAR_PAIR);
            const u32 verdict = lds32(s_lane);
            if (cb && cb->status_seq == KZG_OK && cb->status_lit == KZG_OK && !(nseq && (verdict & 2u))) {
                if (nseq && (verdict & 1u)) {
                    j.lowbits = 0;
                    finish_result(res, j, 0, lit_end, out_end, ovf, rep0, rep1, rep2);
                }
                cb->status_seq = res.status;
                cb->regen = res.regen;
                cb->lit_used = res.lit_used;
                cb->rep_final[0] = res.rep0; cb->rep_final[1] = res.rep1; cb->rep_final[2] = res.rep2;
                cb->bulk_lits = longsum >= KZG_BULK_SUM ? 1u : 0u;
            }
        }
#undef KZG_FSE_WINDOW
        bar_sync(BAR_PAIR);          // queue, ring, tables and hand-over words are free for the next round
    }
#ifndef KZG_FSE_NO_PAIR_RIDE
    // A pair that finds no sequence work left -- 16 384 blocks on 9 472 block slots leave a quarter of the pairs idle in the
    // second wave -- decodes literals too, each of its warps in a piece of the pair's own table area.  (The literal warps
    // were what the launch waited for on log text: ~4.6 tickets per literal warp run as five.)
    {
        const long long deadline = ride_deadline();
        if (deadline) kzg_huf_warp_loop<KZG_HUF_L1_BITS, KZG_HUF_L2CAP, false>(B, reinterpret_cast<u8*>(smem4) + KZG_FSE_CODE_WORDS * 4 + pair * KZG_FSE_PAIR_SMEM + (warp & 1u) * KZG_HUF_WARP_SMEM, lane, deadline, 0xFFFFFFFFu);
    }
#endif
}
