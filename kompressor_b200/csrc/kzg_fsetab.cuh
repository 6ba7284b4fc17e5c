// FSE decode-table build (part of stage s4) for sm_100a: one warp builds the three tables (LL, OF, ML) of
// one compressed block and writes them to the workspace; the sequence kernel (kzg_fse.cuh) only loads them.
// Device-only code.
//
// The textbook build (libzstd: ZSTD_buildFSETable) is two serial passes over the
// table with read-modify-write counters -- a latency chain of ~1 300 dependent shared-memory round trips
// when one thread does it.  Here both passes are data-parallel over the warp:
//   * spread: symbol s occupies the k-th "valid" positions (k in [cum[s], cum[s] + count[s])) of the visiting
//     order pos_j = (j * step) mod size, where positions >= high (reserved for "less than one" symbols) are
//     skipped.  step is odd, so j -> pos_j is a bijection: each lane takes size/32 consecutive j, counts its
//     valid ones, a warp prefix sum gives its first k, and the symbol of placement k comes from an expansion
//     array filled by a running-maximum scan of the symbols' start marks;
//   * numbering: cells are visited in index order 32 at a time; __match_any_sync groups the lanes holding
//     the same symbol, so each lane's `next` value is the symbol's counter plus its rank inside the group.
// Cell format (2 bytes): symbol | next << 6; nbBits and baseline are recomputed by the decoder from `next`.
#pragma once
#include "kzg_stages.cuh"

#define KZG_TAB_WARPS 8
#define KZG_TAB_CELLS 1280                       // per block: LL 512 | ML 512 | OF 256 (u16 each)

namespace kzg_tab {

struct WarpScratch { u16 symk[512]; u16 cells[512]; u16 next[64]; };

__device__ __forceinline__ u32 warp_incl_sum(u32 v, u32 lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= (u32)o) v += t; }
    return v;
}
__device__ __forceinline__ u32 warp_incl_max(u32 v, u32 lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= (u32)o) v = max(v, t); }
    return v;
}

// Builds table t (0 LL, 1 OF, 2 ML) described by d into out[0 .. 2^log).
__device__ __forceinline__ void build(WarpScratch& sc, const KzgFseDesc* d, u32 t, u16* out, u32 lane) {
    if (d->kind[t] == 1) { if (lane == 0) out[0] = (u16)((1u << 6) | (u32)(u16)d->norm[t][0]); return; }
    const u32 log = d->log[t], nsym = d->nsym[t];
    const u32 size = 1u << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    const i16* norm = d->norm[t];
    const i32 c0 = lane < nsym ? (i32)norm[lane] : 0, c1 = lane + 32 < nsym ? (i32)norm[lane + 32] : 0;
    const u32 lt = (1u << lane) - 1u;
    const u32 m0 = __ballot_sync(0xFFFFFFFFu, c0 < 0), m1 = __ballot_sync(0xFFFFFFFFu, c1 < 0);
    const u32 high = size - (u32)__popc(m0) - (u32)__popc(m1);
    const u32 p0 = c0 > 0 ? (u32)c0 : 0u, p1 = c1 > 0 ? (u32)c1 : 0u;
    const u32 in0 = warp_incl_sum(p0, lane), tot0 = __shfl_sync(0xFFFFFFFFu, in0, 31);
    const u32 in1 = warp_incl_sum(p1, lane);
    const u32 ex0 = in0 - p0, ex1 = tot0 + in1 - p1;
    // per-symbol `next` counters start at the count (1 for "less than one" symbols)
    sc.next[lane] = (u16)(c0 < 0 ? 1 : c0);
    sc.next[lane + 32] = (u16)(c1 < 0 ? 1 : c1);
    for (u32 k = lane; k < size; k += 32) sc.symk[k] = 0;
    __syncwarp();
    if (p0) sc.symk[ex0] = (u16)lane;
    if (p1) sc.symk[ex1] = (u16)(lane + 32);
    __syncwarp();
    {   // running maximum: symk[k] = symbol whose placement range contains k
        u32 carry = 0;
        for (u32 c = 0; c < size; c += 32) {
            u32 v = warp_incl_max(sc.symk[c + lane], lane);
            v = max(v, carry);
            sc.symk[c + lane] = (u16)v;
            carry = __shfl_sync(0xFFFFFFFFu, v, 31);
        }
    }
    __syncwarp();
    {   // spread
        const u32 per = size >> 5, j0 = lane * per;          // size >= 32
        u32 cnt = 0;
        for (u32 q = 0; q < per; q++) cnt += (((j0 + q) * step) & mask) < high;
        u32 k = warp_incl_sum(cnt, lane) - cnt;
        for (u32 q = 0; q < per; q++) {
            const u32 pos = ((j0 + q) * step) & mask;
            if (pos < high) { sc.cells[pos] = sc.symk[k]; k++; }
        }
        if (c0 < 0) sc.cells[size - 1 - (u32)__popc(m0 & lt)] = (u16)lane;
        if (c1 < 0) sc.cells[size - 1 - (u32)__popc(m0) - (u32)__popc(m1 & lt)] = (u16)(lane + 32);
    }
    __syncwarp();
    for (u32 c = 0; c < size; c += 32) {     // numbering, in cell order
        const u32 s = sc.cells[c + lane];
        const u32 peers = __match_any_sync(0xFFFFFFFFu, s);
        const u32 rank = (u32)__popc(peers & lt);
        const u32 base = sc.next[s];
        __syncwarp();
        if (rank == 0) sc.next[s] = (u16)(base + (u32)__popc(peers));
        out[c + lane] = (u16)(((base + rank) << 6) | s);
        __syncwarp();
    }
}

}  // namespace kzg_tab

#ifndef KZG_TAB_MINBLOCKS
#define KZG_TAB_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(KZG_TAB_WARPS * 32, KZG_TAB_MINBLOCKS) kzg_fsetab_kernel(KzgBatch B) {
    using namespace kzg_tab;
    __shared__ WarpScratch scratch[KZG_TAB_WARPS];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (B.ctr->overflow & 3u) return;
    const u32 n = B.ctr->n_sblocks;
    WarpScratch& sc = scratch[warp];
    for (u32 si = blockIdx.x * KZG_TAB_WARPS + warp; si < n; si += gridDim.x * KZG_TAB_WARPS) {
        const KzgCBlock* cb = &B.cblocks[B.slist[si]];
        if (cb->status_seq != KZG_OK || cb->status_lit != KZG_OK) continue;
        u16* out = B.fse_tab + (size_t)si * KZG_TAB_CELLS;
        const u32 o0 = cb->fse_owner[0], o1 = cb->fse_owner[1], o2 = cb->fse_owner[2];
        build(sc, kzg_fse_desc_of(B, o0), 0, out, lane);
        __syncwarp();
        build(sc, kzg_fse_desc_of(B, o2), 2, out + 512, lane);
        __syncwarp();
        build(sc, kzg_fse_desc_of(B, o1), 1, out + 1024, lane);
        __syncwarp();
    }
}
