// Bulk stage (s2 and the copy-only parts of s5) for sm_100a: everything whose source and destination are known
// without executing sequences is copied by whole CTAs at full memory-level parallelism, before the execute
// kernel runs (which then skips it).  Device-only code.
//   * raw blocks (copy) and RLE blocks (fill);
//   * compressed blocks without sequences (their literals are the output);
//   * in blocks with sequences: the literals after the last sequence when there are at least KZG_BULK_MIN of
//     them, and every literal run of at least KZG_BULK_MIN bytes inside blocks the sequence kernel flagged
//     (literal-heavy data: a few sequences per block and tens of kilobytes of literals around them).
// One CTA per block of the work list the offsets stage made, taken by ticket.  All copies of a block write
// disjoint ranges, so no synchronisation between them.
#pragma once
#include "kzg_stages.cuh"

#define KZG_BULK_THREADS 256

namespace kzg_bulk {

// CTA-wide copy of n bytes (read-only source, disjoint ranges): 16-byte vector stores, four loads in flight.
__device__ __forceinline__ void cta_copy(u8* dst, const u8* src, u32 n, u32 tid) {
    const u32 head = min(n, (u32)((16 - ((uintptr_t)dst & 15)) & 15));
    if (tid < head) dst[tid] = src[tid];
    dst += head; src += head; n -= head;
    const u32 nvec = n >> 4;
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    if (((uintptr_t)src & 15) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        u32 v = tid;
        for (; v + 3 * KZG_BULK_THREADS < nvec; v += 4 * KZG_BULK_THREADS) {
            const uint4 a = __ldg(s4 + v), b = __ldg(s4 + v + KZG_BULK_THREADS), c = __ldg(s4 + v + 2 * KZG_BULK_THREADS),
                        d = __ldg(s4 + v + 3 * KZG_BULK_THREADS);
            d4[v] = a; d4[v + KZG_BULK_THREADS] = b; d4[v + 2 * KZG_BULK_THREADS] = c; d4[v + 3 * KZG_BULK_THREADS] = d;
        }
        for (; v < nvec; v += KZG_BULK_THREADS) d4[v] = __ldg(s4 + v);
    } else {
        const u32 r = (u32)((uintptr_t)src & 3) * 8;
        const u32* sw = reinterpret_cast<const u32*>(src - ((uintptr_t)src & 3));
        for (u32 v = tid; v < nvec; v += KZG_BULK_THREADS) {
            const u32 a = __ldg(sw + 4 * v), b = __ldg(sw + 4 * v + 1), c = __ldg(sw + 4 * v + 2), d = __ldg(sw + 4 * v + 3);
            const u32 e = r ? __ldg(sw + 4 * v + 4) : 0u;
            d4[v] = make_uint4(__funnelshift_r(a, b, r), __funnelshift_r(b, c, r), __funnelshift_r(c, d, r), __funnelshift_r(d, e, r));
        }
    }
    const u32 done = nvec << 4;
    if (tid < n - done) dst[done + tid] = src[done + tid];
}

__device__ __forceinline__ void cta_fill(u8* dst, u32 byte, u32 n, u32 tid) {
    const u32 head = min(n, (u32)((16 - ((uintptr_t)dst & 15)) & 15));
    if (tid < head) dst[tid] = (u8)byte;
    dst += head; n -= head;
    const u32 w = byte * 0x01010101u, nvec = n >> 4;
    for (u32 v = tid; v < nvec; v += KZG_BULK_THREADS) reinterpret_cast<uint4*>(dst)[v] = make_uint4(w, w, w, w);
    const u32 done = nvec << 4;
    if (tid < n - done) dst[done + tid] = (u8)byte;
}

}  // namespace kzg_bulk

__global__ void __launch_bounds__(KZG_BULK_THREADS) kzg_bulk_kernel(KzgBatch B) {
    using namespace kzg_bulk;
    __shared__ u32 s_ticket;
    __shared__ u32 s_nrun;
    __shared__ u32 s_run[KZG_BULK_THREADS][3];   // long literal runs found in the current chunk of records (at most one per thread)
    const u32 tid = threadIdx.x;
    if (B.ctr->overflow & 3u) return;
    const u32 nwork = B.ctr->n_bulk;
    for (;;) {
        __syncthreads();
        if (tid == 0) { s_ticket = atomicAdd(&B.ctr->ticket_raw, 1u); s_nrun = 0; }
        __syncthreads();
        if (s_ticket >= nwork) break;
        const KzgBlock blk = B.blocks[B.blist[s_ticket]];
        u8* out = B.dst + B.dst_off[blk.frame] + blk.out_off;
        if (blk.type == 0) { cta_copy(out, B.src + blk.src, blk.regen, tid); continue; }
        if (blk.type == 1) { cta_fill(out, B.src[blk.src], blk.regen, tid); continue; }
        const KzgCBlock* cb = &B.cblocks[blk.cidx];
        const u32 lit_type = cb->lit_type, lit_regen = cb->lit_regen, nseq = cb->n_seq;
        const u8* lit = lit_type >= 2 ? B.lit_pool + cb->lit_out : B.src + cb->src + cb->lit_src;
        if (nseq == 0) {
            if (lit_type == 1) cta_fill(out, lit[0], blk.regen, tid);
            else cta_copy(out, lit, blk.regen, tid);
            continue;
        }
        // literals after the last sequence
        const u32 used = cb->lit_used, trail = lit_regen - used;
        if (trail >= KZG_BULK_MIN) {
            if (lit_type == 1) cta_fill(out + (blk.regen - trail), lit[0], trail, tid);
            else cta_copy(out + (blk.regen - trail), lit + used, trail, tid);
        }
        if (!cb->bulk_lits) continue;
        // long literal runs: the threads scan the records, collect the runs, then the CTA copies them one by one
        const u64* seqs = B.seq_pool + cb->seq_out;
        for (u32 i0 = 0; i0 < nseq; i0 += KZG_BULK_THREADS) {
            const u32 i = i0 + tid;
            if (i < nseq) {
                const u64 rec = seqs[i], prev = i ? seqs[i - 1] : 0ull;
                const u32 lit_end = (u32)(rec >> 28) & 0x3FFFFu, p_lit = (u32)(prev >> 28) & 0x3FFFFu, p_out = (u32)(prev >> 46);
                const u32 ll = lit_end - p_lit;
                if (lit_end >= p_lit && ll >= KZG_BULK_MIN && p_out + ll <= blk.regen) {
                    const u32 slot = atomicAdd(&s_nrun, 1u);
                    s_run[slot][0] = p_lit; s_run[slot][1] = p_out; s_run[slot][2] = ll;
                }
            }
            __syncthreads();
            const u32 nrun = s_nrun;
            for (u32 r = 0; r < nrun; r++) {
                if (lit_type == 1) cta_fill(out + s_run[r][1], lit[0], s_run[r][2], tid);
                else cta_copy(out + s_run[r][1], lit + s_run[r][0], s_run[r][2], tid);
            }
            __syncthreads();
            if (tid == 0) s_nrun = 0;
            __syncthreads();
        }
    }
}
