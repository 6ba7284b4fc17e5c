// Literal stage (s3) for sm_100a: Huffman decode-table build + stream decode, one GPU thread per stream,
// eight blocks (8 x 4 streams) per warp.  Device-only code.
//
// Like the sequence kernel this is a serial chain per stream (window -> table entry -> code length ->
// shifted window), so the loop is kept short and free of divergence:
//   * the four lanes of a block build its table together in shared memory.  The table is two-level (kzg_entropy.cuh:
//     512 first-level entries + the few dozen full-table entries of the codes longer than 9 bits, ONE load per symbol
//     either way): 1.1 KiB per block instead of 4 KiB, so an SM holds 128 blocks instead of 48 and the 16 384 blocks of
//     a 1 GiB batch decode in one wave instead of three (the kernel is bound by the latency of the chain, i.e. by how
//     many chains are in flight).  Blocks with long streams (>= 32 KiB of literals: literal-heavy data, where few
//     chains exist anyway and every cycle of the chain counts) and blocks whose codes need more second-level entries
//     than the compact layout holds (never seen outside crafted input) are left to a second launch with full
//     single-level tables;
//   * each lane's backward bitstream is staged by cp.async (LDGSTS: no destination register, no per-warp
//     scoreboard stall) into a 128-byte ring, serviced at warp-uniform points once per 16 symbols;
//   * a 64-bit window is re-read from the ring every 4 symbols (3 LDS + 2 funnel shifts, 4 x 11 bits fit),
//     symbols are packed 16 at a time into one 16-byte store;
//   * lanes that run out of symbols keep executing on frozen positions (stores and bit consumption masked).
// What libzstd does here is HUF_decompress4X1 / HUF_readDTableX1, reached from the reference through
// ZSTD_decompressStream (kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:178).
#pragma once
#include "kzg_stages.cuh"

#define KZG_HUF_WARPS 2
#ifndef KZG_HUF_INFLIGHT
#define KZG_HUF_INFLIGHT 3                                   // copy groups (one per ring service) that may still be in flight at a service point: 1..3
#endif
#define KZG_HUF_L2CAP 64u                                    // second-level entries per block, compact layout
#define KZG_HUF_LONG_STREAMS 32768u                          // blocks with at least this many literals go to the full-table launch
#define KZG_HUF_RING_BYTES (32u * 128u)                      // per warp: 32 lanes x 8 chunks x 16 B
#define KZG_HUF_WARP_SMEM_OF(l1bits, l2cap) (8u * ((2u << (l1bits)) + 2u * (l2cap)) + KZG_HUF_RING_BYTES)
#define KZG_HUF_WARP_SMEM KZG_HUF_WARP_SMEM_OF(KZG_HUF_L1_BITS, KZG_HUF_L2CAP)
#define KZG_HUF_SMEM (KZG_HUF_WARPS * KZG_HUF_WARP_SMEM)
#define KZG_HUF_SMEM_FULL (KZG_HUF_WARPS * KZG_HUF_WARP_SMEM_OF(KZG_HUF_MAXBITS, 0u))

static_assert(sizeof(KzgHufScratch) <= 4 * 128, "table-build scratch must fit the rings of a block's four lanes");

namespace kzg_huf {

__device__ __forceinline__ u32 lds16(u32 a) { u16 v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void cp_async16(u32 dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
// Predicated form: one instruction, no branch around it.
__device__ __forceinline__ void cp_async16_if(bool p, u32 dst, const void* src) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" :: "r"(dst), "l"(src), "r"((u32)p) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

}  // namespace kzg_huf

// The work loop of one warp (wbase: its KZG_HUF_WARP_SMEM bytes of shared memory).  With `deadline` set (a clock64()
// value), the warp takes no new ticket after it: that is how the warps that ride along in the sequence kernel
// (kzg_fse.cuh) leave when the sequence work is about to end; kzg_huf_kernel then finishes what is left.  `long_min`:
// blocks with at least this many literals are left to the full-table launch (riders take them all).
// L1BITS / L2CAP: first-level index width and second-level entries per block in this warp's layout; SECOND: the launch
// that takes the blocks the compact launch left (work list hlist2), with L1BITS = KZG_HUF_MAXBITS (full tables).
template <u32 L1BITS, u32 L2CAP, bool SECOND>
__device__ __forceinline__ void kzg_huf_warp_loop(const KzgBatch& B, u8* wbase, u32 lane, long long deadline, u32 long_min) {
    using namespace kzg_huf;
    const u32 hb = lane >> 2, st = lane & 3;
    constexpr u32 L1_BYTES = 2u << L1BITS;
    u16* l1 = reinterpret_cast<u16*>(wbase + hb * L1_BYTES);
    u16* l2 = reinterpret_cast<u16*>(wbase + 8u * L1_BYTES + hb * (2u * L2CAP));
    u8* const rings = wbase + 8u * (L1_BYTES + 2u * L2CAP);
    KzgHufScratch* sc = reinterpret_cast<KzgHufScratch*>(rings + hb * 512u);   // = the block's four rings
    const u32 s_l1 = (u32)__cvta_generic_to_shared(l1), s_l2 = (u32)__cvta_generic_to_shared(l2);
    const u32 s_ring = (u32)__cvta_generic_to_shared(rings) + (lane << 7);
    if (B.ctr->overflow & 3u) return;
    const u32 n = SECOND ? B.ctr->n_hblocks2 : B.ctr->n_hblocks;
    const u32* list = SECOND ? B.hlist2 : B.hlist;
    for (;;) {
        u32 g = 0xFFFFFFFFu;
        if (lane == 0 && !(deadline && clock64() > deadline)) g = atomicAdd(SECOND ? &B.ctr->ticket_huf2 : &B.ctr->ticket_huf, 1u);
        g = __shfl_sync(0xFFFFFFFFu, g, 0);
        if (g == 0xFFFFFFFFu || (u64)g * 8 >= n) break;
        const u32 h = g * 8 + hb;
        bool valid = h < n;
        KzgCBlock* cb = nullptr;
        if (valid) {
            cb = &B.cblocks[list[h]];
            valid = cb->status_lit == KZG_OK && cb->huf_owner != KZG_INVALID_OWNER;
            if (valid && !SECOND && cb->lit_regen >= long_min) {      // long streams: left to the full-table launch
                if (st == 0) B.hlist2[atomicAdd(&B.ctr->n_hblocks2, 1u)] = list[h];
                valid = false;
            }
        }
        // ---- table: built by the block's four lanes (the rings serve as scratch until the streams start)
        u32 maxbits = KZG_HUF_MAXBITS, nsym = 0, nlong = 0;
        if (valid) {
            const KzgHufDesc* hd = kzg_huf_desc_of(B, cb->huf_owner);
            maxbits = hd->maxbits; nsym = hd->nsym;
            kzg_huf_build_phase1(sc, hd->weight, nsym, st);
        }
        __syncwarp();
        if (valid) {
            nlong = kzg_huf_long_entries(sc, maxbits, L1BITS);
            if (nlong > L2CAP) {                // more long codes than the compact layout holds: left to the full-table launch
                if (!SECOND && st == 0) B.hlist2[atomicAdd(&B.ctr->n_hblocks2, 1u)] = list[h];
                valid = false;
            }
        }
        if (valid) kzg_huf_build_phase2(sc, maxbits, st);
        __syncwarp();
        if (valid) kzg_huf_build_phase3(l1, l2, sc, nsym, maxbits, st, L1BITS);
        __syncwarp();

        // ---- this lane's stream
        const u8* base = nullptr;
        u8* dst = nullptr;
        u32 cnt = 0;
        // (lanes without a stream run the loops on a frozen position: one ring word each, or they all hit the same bank)
        i32 P = (i32)(lane << 5), lowbits = 0, next_c = -1, ready_lo = 0, pend1 = 0, pend2 = 0, pend3 = 0, pend4 = 0;
        bool mine = valid && st < cb->n_streams, bad = false;
        if (mine) {
            const u32 regen = cb->lit_regen;
            const u8* src = B.src + cb->src + cb->lit_src;
            dst = B.lit_pool + cb->lit_out;
            cnt = regen;
            u32 size = cb->stream_size[0];
            if (cb->n_streams == 4) {
                const u32 seg = (regen + 3) >> 2;
                for (u32 k = 0; k < st; k++) src += cb->stream_size[k];
                size = cb->stream_size[st];
                dst += st * seg;
                cnt = st < 3 ? seg : regen - 3 * seg;
            }
            const u32 last = size ? src[size - 1] : 0u;
            if (last == 0) { bad = true; cnt = 0; }
            else {
                base = (const u8*)((uintptr_t)src & ~(uintptr_t)15);
                lowbits = (i32)(8u * (u32)(src - base));
                P = lowbits + (i32)(8u * (size - 1)) + kzg_highbit(last);
                next_c = P >> 7;
                for (int r = 0; r < 8; r++)
                    if (next_c >= 0) { cp_async16(s_ring + (((u32)next_c & 7u) << 4), base + ((size_t)next_c << 4)); next_c--; }
            }
        }
        cp_commit();
        cp_wait<0>();
        ready_lo = pend1 = pend2 = pend3 = pend4 = next_c + 1;
        const u32 sh = 32u - maxbits, sshift = kzg_huf_split(maxbits, L1BITS);

#define KZG_HUF_WINDOW(hi, lo)                                                     \
        {                                                                          \
            const u32 t_ = (u32)(P >> 3), q_ = (u32)P & 31u;                       \
            const u32 w2_ = lds32(s_ring + (t_ & 0x7Cu));                          \
            const u32 w1_ = lds32(s_ring + ((t_ - 4u) & 0x7Cu));                   \
            const u32 w0_ = lds32(s_ring + ((t_ - 8u) & 0x7Cu));                   \
            hi = __funnelshift_r(w1_, w2_, q_);                                    \
            lo = __funnelshift_r(w0_, w1_, q_);                                    \
        }
#define KZG_HUF_SERVICE(need_bits)                                                                        \
        {                                                                                                 \
            cp_wait<KZG_HUF_INFLIGHT>();                                                                  \
            ready_lo = KZG_HUF_INFLIGHT == 1 ? pend2 : (KZG_HUF_INFLIGHT == 2 ? pend3 : pend4);           \
            const i32 cP_ = P >> 7;                                                                       \
            const i32 want_ = min(min(next_c - (cP_ - 8), next_c + 1), 2);                                \
            const u8* g_ = base + ((i64)next_c << 4);                                                     \
            cp_async16_if(want_ > 0, s_ring + (((u32)next_c & 7u) << 4), g_);                             \
            cp_async16_if(want_ > 1, s_ring + (((u32)(next_c - 1) & 7u) << 4), g_ - 16);                  \
            next_c -= max(want_, 0);                                                                      \
            cp_commit();                                                                                  \
            pend4 = pend3;                                                                                \
            pend3 = pend2;                                                                                \
            pend2 = pend1;                                                                                \
            pend1 = next_c + 1;                                                                           \
            const i32 need_lo_ = max((P - (i32)(need_bits)) >> 7, 0);                                     \
            if (need_lo_ < ready_lo) { cp_wait<0>(); ready_lo = pend4 = pend3 = pend2 = pend1; }          \
        }
        // one symbol from the top of hi:lo; advances the window and `used`
#define KZG_HUF_SYM(sym, hi, lo, used)                                             \
        {                                                                          \
            const u32 i_ = (hi) >> sh;                                             \
            const u32 e_ = lds16(L1BITS >= KZG_HUF_MAXBITS ? s_l1 + (i_ << 1) : (i_ < nlong ? s_l2 + (i_ << 1) : s_l1 + ((i_ >> sshift) << 1))); \
            const u32 l_ = e_ >> 8;                                                \
            sym = e_ & 255u;                                                       \
            hi = __funnelshift_l(lo, hi, l_);                                      \
            lo <<= l_;                                                             \
            used += l_;                                                            \
        }

        // ---- head: single symbols until the destination is 16-byte aligned (at most 15)
        u32 done = 0;
        {
            const u32 hcnt = min(cnt, (u32)((0 - (uintptr_t)dst) & 15));
            const u32 hmax = __reduce_max_sync(0xFFFFFFFFu, hcnt);
            for (u32 k = 0; k < hmax; k++) {
                u32 hi, lo, s, used = 0;
                KZG_HUF_WINDOW(hi, lo);
                KZG_HUF_SYM(s, hi, lo, used);
                if (k < hcnt) { dst[k] = (u8)s; P -= (i32)used; }
            }
            done = hcnt;
        }
        // ---- body: 16 symbols per 16-byte store
        {
            const u32 groups = (cnt - done) >> 4;
            const u32 gmax = __reduce_max_sync(0xFFFFFFFFu, groups);
            uint4* d4 = reinterpret_cast<uint4*>(dst + done);
            for (u32 gi = 0; gi < gmax; gi++) {
                KZG_HUF_SERVICE(16 * 11 + 64 + 31);
                const bool act = gi < groups;
                u32 v[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    u32 hi, lo, s0, s1, s2, s3, used = 0;
                    KZG_HUF_WINDOW(hi, lo);
                    KZG_HUF_SYM(s0, hi, lo, used);
                    KZG_HUF_SYM(s1, hi, lo, used);
                    KZG_HUF_SYM(s2, hi, lo, used);
                    KZG_HUF_SYM(s3, hi, lo, used);
                    v[q] = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
                    P -= (i32)(act ? used : 0u);
                }
                if (act) d4[gi] = make_uint4(v[0], v[1], v[2], v[3]);
            }
            done += groups << 4;
        }
        // ---- tail: fewer than 16 symbols
        {
            KZG_HUF_SERVICE(15 * 11 + 64 + 31);
            const u32 tcnt = cnt - done;
            const u32 tmax = __reduce_max_sync(0xFFFFFFFFu, tcnt);
            for (u32 k = 0; k < tmax; k++) {
                u32 hi, lo, s, used = 0;
                KZG_HUF_WINDOW(hi, lo);
                KZG_HUF_SYM(s, hi, lo, used);
                if (k < tcnt) { dst[done + k] = (u8)s; P -= (i32)used; }
            }
        }
#undef KZG_HUF_SYM
#undef KZG_HUF_SERVICE
#undef KZG_HUF_WINDOW
        cp_wait<0>();
        // a stream that is not consumed exactly: corrupt, or (libzstd's fast loops) accepted / flagged -- kzg_huf_verdict
        if (mine && (bad || kzg_huf_verdict(cb, P == lowbits ? 0u : (P > lowbits ? 1u : 2u)))) cb->status_lit = KZG_E_CORRUPTION;
        __syncwarp();
    }
}

template <u32 L1BITS, u32 L2CAP, bool SECOND>
__global__ void __launch_bounds__(KZG_HUF_WARPS * 32) kzg_huf_kernel(KzgBatch B) {
    extern __shared__ uint4 smem4[];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    kzg_huf_warp_loop<L1BITS, L2CAP, SECOND>(B, reinterpret_cast<u8*>(smem4) + warp * KZG_HUF_WARP_SMEM_OF(L1BITS, L2CAP), lane, 0ll, B.huf_long_min);
}
