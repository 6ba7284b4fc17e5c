// B200 (sm_100a) kernels of the batched Zstandard decoder.
//
// Pipeline (one stream, no host round trip until the per-frame status is read back):
//   scan      thread / frame        frame header, block-header walk -> block work list       (stage s1)
//   parse     thread / block        literals + sequences section headers, Huffman weights,
//                                   FSE normalised counts, pool allocation                   (s3/s4 headers)
//   link      thread / frame        treeless / repeat-mode table owners inside a frame
//   huffman   thread / stream       table build (4 lanes per table) + 4-stream literal decode (s3)
//   fse       thread / block        table build + sequential ZSTD_decodeSequence state machine (s4)
//   offsets   thread / frame        per-block output offsets, repeat-offset hand-over, size checks
//   execute   CTA / frame           raw + RLE blocks, literal + match copy through a shared-memory
//                                   output tile, 16-byte vectorised flush                    (s2, s5)
//   finish    4 lanes / frame       XXH64 of the output, checksum compare, status + size      (s6)
//
// The format logic follows RFC 8878 / SURVEY.md Appendix A; it replaces what libzstd does underneath
// the reference call ZSTD_decompressStream (kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:178).
#include "kzg_kernels.cuh"
#include "kzg_stages.cuh"
#include "kzg_fse.cuh"

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) kzg_scan_kernel(KzgBatch B) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B.n_frames) kzg_stage_scan(B, i);
}

__global__ void __launch_bounds__(64) kzg_parse_kernel(KzgBatch B) {
    if (B.ctr->overflow & 3u) return;    // block tables overflowed: the host grows them and re-runs
    const u32 n = min(B.ctr->n_cblocks, B.cblock_cap);
    for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) kzg_stage_parse(B, c);
}

__global__ void __launch_bounds__(128) kzg_link_kernel(KzgBatch B) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.n_frames || (B.ctr->overflow & 3u)) return;
    kzg_stage_link(B, i);
}

// ------------------------------------------------------------------------------------------------
// huffman: one warp decodes 8 blocks at a time = 32 streams, one thread per stream. The four threads
// of a block build its 2^maxbits-entry table together in shared memory (4 KiB + 512 B scratch per block).
#define KZG_HUF_WARPS 2
#define KZG_HUF_TABLE_BYTES ((1u << KZG_HUF_MAXBITS) * 2u)
#define KZG_HUF_SMEM (KZG_HUF_WARPS * 8 * (KZG_HUF_TABLE_BYTES + sizeof(KzgHufScratch)))
__global__ void __launch_bounds__(KZG_HUF_WARPS * 32) kzg_huf_kernel(KzgBatch B) {
    extern __shared__ uint4 smem4[];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 hb = lane >> 2, st = lane & 3;
    u16* tab = reinterpret_cast<u16*>(smem4) + (size_t)(warp * 8 + hb) * (KZG_HUF_TABLE_BYTES / 2);
    KzgHufScratch* sc = reinterpret_cast<KzgHufScratch*>(reinterpret_cast<u8*>(smem4) + KZG_HUF_WARPS * 8 * KZG_HUF_TABLE_BYTES) + (warp * 8 + hb);
    if (B.ctr->overflow & 3u) return;
    const u32 n = B.ctr->n_hblocks;
    for (;;) {
        u32 g = 0;
        if (lane == 0) g = atomicAdd(&B.ctr->ticket_huf, 1u);
        g = __shfl_sync(0xFFFFFFFFu, g, 0);
        if (g * 8 >= n) break;
        const u32 h = g * 8 + hb;
        bool valid = h < n;
        KzgCBlock* cb = nullptr;
        if (valid) {
            cb = &B.cblocks[B.hlist[h]];
            valid = cb->status_lit == KZG_OK && cb->huf_owner != KZG_INVALID_OWNER;
        }
        u32 maxbits = 0, nsym = 0;
        if (valid) {
            const KzgHufDesc* hd = &B.huf[cb->huf_owner];
            maxbits = hd->maxbits; nsym = hd->nsym;
            kzg_huf_build_phase1(sc, hd->weight, nsym, st);
        }
        __syncwarp();
        if (valid) kzg_huf_build_phase2(sc, maxbits, st);
        __syncwarp();
        if (valid) kzg_huf_build_phase3(tab, sc, nsym, maxbits, st);
        __syncwarp();
        if (valid && st < cb->n_streams) kzg_stage_huf_stream(B, cb, st, tab, maxbits);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// fse: see kzg_fse.cuh (one thread per block, lane-interleaved tables, cp.async bitstream ring)
__global__ void __launch_bounds__(128) kzg_offsets_kernel(KzgBatch B) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.n_frames || (B.ctr->overflow & 3u)) return;
    kzg_stage_offsets(B, i);
}

// ------------------------------------------------------------------------------------------------
// execute: one CTA per frame
#define KZG_EXEC_THREADS 256
#define KZG_EXEC_WARPS 8
#define KZG_LONG_ML 48u
#define KZG_LONG_LL 16u
#define KZG_TILE_SLACK 64u

struct KzgExecShared { u32 frame; u32 err; volatile u32 done; volatile u32 P; };

__device__ __forceinline__ void kzg_copy_global(u8* dst, const u8* src, u32 n, u32 tid, u32 nthreads) {
    const u32 head = min(n, (u32)((16 - ((uintptr_t)dst & 15)) & 15));
    for (u32 k = tid; k < head; k += nthreads) dst[k] = src[k];
    dst += head; src += head; n -= head;
    const u32 nvec = n >> 4;
    if (((uintptr_t)src & 15) == 0) {
        for (u32 v = tid; v < nvec; v += nthreads)
            reinterpret_cast<uint4*>(dst)[v] = reinterpret_cast<const uint4*>(src)[v];
    } else {
        const u32 r = (u32)((uintptr_t)src & 3) * 8;
        const u32* sw = reinterpret_cast<const u32*>(src - ((uintptr_t)src & 3));
        for (u32 v = tid; v < nvec; v += nthreads) {
            const u32 a = sw[4 * v], b = sw[4 * v + 1], c = sw[4 * v + 2], d = sw[4 * v + 3];
            const u32 e = r ? sw[4 * v + 4] : 0u;
            reinterpret_cast<uint4*>(dst)[v] = make_uint4(__funnelshift_r(a, b, r), __funnelshift_r(b, c, r),
                                                          __funnelshift_r(c, d, r), __funnelshift_r(d, e, r));
        }
    }
    for (u32 k = (nvec << 4) + tid; k < n; k += nthreads) dst[k] = src[k];
}

__device__ __forceinline__ void kzg_set_global(u8* dst, u32 byte, u32 n, u32 tid, u32 nthreads) {
    const u32 head = min(n, (u32)((16 - ((uintptr_t)dst & 15)) & 15));
    for (u32 k = tid; k < head; k += nthreads) dst[k] = (u8)byte;
    dst += head; n -= head;
    const u32 w = byte * 0x01010101u, nvec = n >> 4;
    for (u32 v = tid; v < nvec; v += nthreads) reinterpret_cast<uint4*>(dst)[v] = make_uint4(w, w, w, w);
    for (u32 k = (nvec << 4) + tid; k < n; k += nthreads) dst[k] = (u8)byte;
}

// One lane copies one match of <= KZG_LONG_ML bytes inside the tile. spos may be negative: that part
// of the source lies in earlier blocks of the frame (already in global memory at gout[spos]).
// Byte at block-relative position p < 0: earlier output of the frame, or (before the frame start) the
// raw-content dictionary that prefixes every frame of the batch.
__device__ __forceinline__ u8 kzg_hist_byte(const u8* gout, i32 p, i64 boff, const u8* dict_end) {
    const i64 fi = boff + p;
    return fi >= 0 ? gout[p] : dict_end[fi];
}

__device__ __forceinline__ void kzg_copy_match_lane(u8* tile, const u8* gout, u32 mdst, i32 spos, u32 ml, u32 off,
                                                    i64 boff, const u8* dict_end) {
    while (spos < 0 && ml) { tile[mdst] = kzg_hist_byte(gout, spos, boff, dict_end); mdst++; spos++; ml--; }
    if (!ml) return;
    u8* q = tile + mdst;
    const u8* s = tile + spos;
    if (off < 8) { for (u32 k = 0; k < ml; k++) q[k] = s[k]; return; }
    while (((uintptr_t)q & 3) && ml) { *q++ = *s++; ml--; }
    const u32 nw = ml >> 2;
    if (nw) {
        const u32 r = (u32)((uintptr_t)s & 3);
        const u32* sw = reinterpret_cast<const u32*>(s - r);
        u32* qw = reinterpret_cast<u32*>(q);
        u32 lo = sw[0];
        for (u32 k = 0; k < nw; k++) {
            const u32 hi = sw[k + 1];
            qw[k] = __funnelshift_r(lo, hi, r * 8);
            lo = hi;
        }
    }
    for (u32 k = nw << 2; k < ml; k++) q[k] = s[k];
}

// Returns true (uniformly) if the block is corrupt.
__device__ bool kzg_exec_cblock(const KzgBatch& B, const KzgBlock& blk, u8* out, u8* tile0, KzgExecShared& sh,
                                u64 hist_before_block) {
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const KzgCBlock* cb = &B.cblocks[blk.cidx];
    const u32 nseq = cb->n_seq, regen = blk.regen, lit_regen = cb->lit_regen, lit_type = cb->lit_type;
    const u8* lit = lit_type >= 2 ? B.lit_pool + cb->lit_out : B.src + cb->src + cb->lit_src;
    const u32 rle = lit_type == 1 ? (u32)lit[0] : 0u;
    if (nseq == 0) {
        if (lit_type == 1) kzg_set_global(out, rle, regen, tid, KZG_EXEC_THREADS);
        else kzg_copy_global(out, lit, regen, tid, KZG_EXEC_THREADS);
        return false;
    }
    if (regen + KZG_TILE_SLACK > B.tile_bytes) {      // host re-runs with a larger tile
        if (tid == 0) atomicOr(&B.ctr->overflow, 16u);
        return true;
    }
    u8* tile = tile0 + ((uintptr_t)out & 15);
    const u64* seqs = B.seq_pool + cb->seq_out;
    const u32 nb = (nseq + 31) >> 5;
    const u32 ri0 = cb->rep_init[0], ri1 = cb->rep_init[1], ri2 = cb->rep_init[2];
    const i64 boff = (i64)blk.out_off;
    const u8* dict_end = B.dict + B.dict_size;
    if (tid == 0) { sh.done = 0; sh.P = 0; sh.err = 0; }

    // ---- phase A: all literal runs (no dependencies)
    for (u32 b = warp; b < nb; b += KZG_EXEC_WARPS) {
        const u32 i = b * 32 + lane;
        const u64 rec = i < nseq ? seqs[i] : 0ull;
        u64 prev = __shfl_up_sync(0xFFFFFFFFu, rec, 1);
        if (lane == 0) prev = b ? seqs[i - 1] : 0ull;
        const u32 lit_end = (u32)(rec >> 28) & 0x3FFFFu, p_lit_end = (u32)(prev >> 28) & 0x3FFFFu;
        const u32 d = (u32)(prev >> 46);
        const u32 ll = i < nseq ? lit_end - p_lit_end : 0u;
        const u32 n1 = min(ll, KZG_LONG_LL);
        if (lit_type == 1) { for (u32 k = 0; k < n1; k++) tile[d + k] = (u8)rle; }
        else { for (u32 k = 0; k < n1; k++) tile[d + k] = lit[p_lit_end + k]; }
        u32 m = __ballot_sync(0xFFFFFFFFu, ll > KZG_LONG_LL);
        while (m) {
            const u32 sl = __ffs(m) - 1;
            m &= m - 1;
            const u32 dd = __shfl_sync(0xFFFFFFFFu, d, sl), lx = __shfl_sync(0xFFFFFFFFu, p_lit_end, sl),
                      nn = __shfl_sync(0xFFFFFFFFu, ll, sl);
            for (u32 k = KZG_LONG_LL + lane; k < nn; k += 32) tile[dd + k] = lit_type == 1 ? (u8)rle : lit[lx + k];
        }
    }
    {   // literals after the last sequence
        const u64 lastrec = seqs[nseq - 1];
        const u32 used = (u32)(lastrec >> 28) & 0x3FFFFu, oe = (u32)(lastrec >> 46);
        for (u32 k = tid; k < lit_regen - used; k += KZG_EXEC_THREADS) tile[oe + k] = lit_type == 1 ? (u8)rle : lit[used + k];
    }
    __syncthreads();

    // ---- phase B: matches. Batches of 32 sequences are taken round-robin by the warps and published
    // in order; sh.P is the output position below which every byte of the block is final.
    for (u32 b = warp; b < nb; b += KZG_EXEC_WARPS) {
        const u32 i = b * 32 + lane;
        const bool valid = i < nseq;
        const u64 rec = valid ? seqs[i] : 0ull;
        u64 prev = __shfl_up_sync(0xFFFFFFFFu, rec, 1);
        if (lane == 0) prev = b ? seqs[i - 1] : 0ull;
        const u32 lit_end = (u32)(rec >> 28) & 0x3FFFFu, p_lit_end = (u32)(prev >> 28) & 0x3FFFFu;
        const u32 out_end = (u32)(rec >> 46), p_out_end = (u32)(prev >> 46);
        const u32 ll = lit_end - p_lit_end;
        const u32 mdst = p_out_end + ll;
        const u32 ml = valid ? out_end - mdst : 0u;
        u32 off = (u32)rec & KZG_OFF_MASK;
        if (off & KZG_OFF_SYM) {
            const u32 k = off & 3, dd = (off & (KZG_OFF_SYM - 1)) >> 2;
            const u32 base = k == 0 ? ri0 : (k == 1 ? ri1 : ri2);
            off = (base >= KZG_OFF_INVALID || base <= dd) ? KZG_OFF_INVALID : base - dd;
        }
        const bool bad = valid && (off == 0 || off >= KZG_OFF_INVALID || (u64)off > hist_before_block + mdst);
        if (bad) atomicOr(&sh.err, 1u);
        bool active = valid && !bad && ml > 0;
        const i32 spos = (i32)mdst - (i32)off;
        u32 need = mdst;                                         // overlapping match: needs everything below itself
        if (off >= ml) { const i32 e = spos + (i32)ml; need = e > 0 ? (u32)e : 0u; }
        const u32 lastlane = min(31u, nseq - 1 - b * 32);
        const u32 batch_end = __shfl_sync(0xFFFFFFFFu, out_end, lastlane);
        for (;;) {
            const u32 am = __ballot_sync(0xFFFFFFFFu, active);
            if (!am) break;
            u32 P = 0, oldest = 0;
            if (lane == 0) { P = sh.P; oldest = (sh.done == b); }
            P = __shfl_sync(0xFFFFFFFFu, P, 0);
            oldest = __shfl_sync(0xFFFFFFFFu, oldest, 0);
            u32 hw = P;
            if (oldest) hw = __shfl_sync(0xFFFFFFFFu, mdst, __ffs(am) - 1);   // all below the first pending match is final
            const bool ready = active && need <= hw;
            const u32 rm = __ballot_sync(0xFFFFFFFFu, ready);
            if (!rm) { __nanosleep(64); continue; }
            if (ready && ml <= KZG_LONG_ML) kzg_copy_match_lane(tile, out, mdst, spos, ml, off, boff, dict_end);
            u32 lm = __ballot_sync(0xFFFFFFFFu, ready && ml > KZG_LONG_ML);
            while (lm) {
                const u32 sl = __ffs(lm) - 1;
                lm &= lm - 1;
                const u32 md = __shfl_sync(0xFFFFFFFFu, mdst, sl), nn = __shfl_sync(0xFFFFFFFFu, ml, sl),
                          of = __shfl_sync(0xFFFFFFFFu, off, sl);
                const i32 sp = __shfl_sync(0xFFFFFFFFu, spos, sl);
                if (of >= nn) {
                    for (u32 k = lane; k < nn; k += 32) { const i32 p = sp + (i32)k; tile[md + k] = p < 0 ? kzg_hist_byte(out, p, boff, dict_end) : tile[p]; }
                } else {   // overlapping: every byte equals one of the `of` bytes just before the match
                    for (u32 k = lane; k < nn; k += 32) { const i32 p = sp + (i32)(k % of); tile[md + k] = p < 0 ? kzg_hist_byte(out, p, boff, dict_end) : tile[p]; }
                }
            }
            if (ready) active = false;
            __syncwarp();
        }
        if (lane == 0) { while (sh.done != b) __nanosleep(32); }
        __syncwarp();
        __threadfence_block();
        if (lane == 0) { sh.P = batch_end; __threadfence_block(); sh.done = b + 1; }
    }
    __syncthreads();

    // ---- flush the tile: tile and out share their alignment modulo 16
    {
        const u32 head = min(regen, (u32)((16 - ((uintptr_t)out & 15)) & 15));
        for (u32 k = tid; k < head; k += KZG_EXEC_THREADS) out[k] = tile[k];
        const u32 nvec = (regen - head) >> 4;
        const uint4* tv = reinterpret_cast<const uint4*>(tile + head);
        uint4* ov = reinterpret_cast<uint4*>(out + head);
        for (u32 v = tid; v < nvec; v += KZG_EXEC_THREADS) ov[v] = tv[v];
        for (u32 k = head + (nvec << 4) + tid; k < regen; k += KZG_EXEC_THREADS) out[k] = tile[k];
    }
    return sh.err != 0;
}

__global__ void __launch_bounds__(KZG_EXEC_THREADS) kzg_exec_kernel(KzgBatch B) {
    extern __shared__ uint4 smem4[];
    u8* tile0 = reinterpret_cast<u8*>(smem4);
    __shared__ KzgExecShared sh;
    const u32 tid = threadIdx.x;
    if (B.ctr->overflow & 3u) return;
    for (;;) {
        __syncthreads();
        if (tid == 0) sh.frame = atomicAdd(&B.ctr->ticket_exec, 1u);
        __syncthreads();
        const u32 f = sh.frame;
        if (f >= B.n_frames) break;
        const KzgFrame* fr = &B.frames[f];
        if (fr->status != KZG_OK) continue;
        u8* const fdst = B.dst + B.dst_off[f];
        const u32 nblk = fr->n_blocks, first = fr->first_block;
        bool failed = false;
        for (u32 k = 0; k < nblk && !failed; k++) {
            const KzgBlock blk = B.blocks[first + k];
            u8* out = fdst + blk.out_off;
            if (blk.type == 0) kzg_copy_global(out, B.src + blk.src, blk.regen, tid, KZG_EXEC_THREADS);
            else if (blk.type == 1) kzg_set_global(out, B.src[blk.src], blk.regen, tid, KZG_EXEC_THREADS);
            else failed = kzg_exec_cblock(B, blk, out, tile0, sh, blk.out_off + B.dict_size);
            __syncthreads();
        }
        if (failed && tid == 0) B.frames[f].status = (B.ctr->overflow & 16u) ? KZG_E_INTERNAL_OVERFLOW : KZG_E_CORRUPTION;
    }
}

// ------------------------------------------------------------------------------------------------
// finish: XXH64 (Appendix A.8) with the four accumulator lanes of a frame on four threads, then the
// checksum compare and the per-frame status / size. 8 frames per warp.
#define XP1 0x9E3779B185EBCA87ull
#define XP2 0xC2B2AE3D27D4EB4Full
#define XP3 0x165667B19E3779F9ull
#define XP4 0x85EBCA77C2B2AE63ull
#define XP5 0x27D4EB2F165667C5ull
__device__ __forceinline__ u64 kzg_rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ u64 kzg_xxround(u64 acc, u64 x) { return kzg_rotl64(acc + x * XP2, 31) * XP1; }
__device__ __forceinline__ u64 kzg_ld64_any(const u8* p) {
    if (((uintptr_t)p & 7) == 0) return *reinterpret_cast<const u64*>(p);
    if (((uintptr_t)p & 3) == 0) { const u32* q = reinterpret_cast<const u32*>(p); return (u64)q[0] | ((u64)q[1] << 32); }
    return kzg_rd64(p);
}

__global__ void __launch_bounds__(128) kzg_finish_kernel(KzgBatch B) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 f = t >> 2, q = t & 3;
    const bool in_range = f < B.n_frames;
    u32 status = KZG_OK;
    u64 len = 0;
    bool do_sum = false;
    const KzgFrame* fr = nullptr;
    if (in_range) {
        fr = &B.frames[f];
        status = (B.ctr->overflow & 3u) ? (u32)KZG_E_INTERNAL_OVERFLOW : fr->status;
        len = fr->out_size;
        do_sum = status == KZG_OK && fr->checksum_off != KZG_NONE;
    }
    u64 acc = 0;
    const u8* p = nullptr;
    if (do_sum) {
        p = B.dst + B.dst_off[f];
        acc = q == 0 ? XP1 + XP2 : (q == 1 ? XP2 : (q == 2 ? 0ull : 0ull - XP1));
        const u64 stripes = len >> 5;
        const u8* s = p + 8 * q;
        u64 k = 0;
        for (; k + 4 <= stripes; k += 4) {
            const u64 x0 = kzg_ld64_any(s + 32 * k), x1 = kzg_ld64_any(s + 32 * (k + 1)),
                      x2 = kzg_ld64_any(s + 32 * (k + 2)), x3 = kzg_ld64_any(s + 32 * (k + 3));
            acc = kzg_xxround(acc, x0); acc = kzg_xxround(acc, x1);
            acc = kzg_xxround(acc, x2); acc = kzg_xxround(acc, x3);
        }
        for (; k < stripes; k++) acc = kzg_xxround(acc, kzg_ld64_any(s + 32 * k));
    }
    // gather the four lanes on lane q == 0 (all 32 threads take part in the shuffles)
    const u32 base_lane = (threadIdx.x & 31) & ~3u;
    const u64 v2 = __shfl_sync(0xFFFFFFFFu, acc, base_lane + 1);
    const u64 v3 = __shfl_sync(0xFFFFFFFFu, acc, base_lane + 2);
    const u64 v4 = __shfl_sync(0xFFFFFFFFu, acc, base_lane + 3);
    if (!in_range || q != 0) return;
    if (do_sum) {
        u64 h;
        if (len >= 32) {
            h = kzg_rotl64(acc, 1) + kzg_rotl64(v2, 7) + kzg_rotl64(v3, 12) + kzg_rotl64(v4, 18);
            h = (h ^ kzg_xxround(0, acc)) * XP1 + XP4;
            h = (h ^ kzg_xxround(0, v2)) * XP1 + XP4;
            h = (h ^ kzg_xxround(0, v3)) * XP1 + XP4;
            h = (h ^ kzg_xxround(0, v4)) * XP1 + XP4;
        } else h = XP5;
        h += len;
        const u8* e = p + len;
        const u8* c = p + (len & ~31ull);
        while (e - c >= 8) { h = kzg_rotl64(h ^ kzg_xxround(0, kzg_ld64_any(c)), 27) * XP1 + XP4; c += 8; }
        if (e - c >= 4) { h = kzg_rotl64(h ^ ((u64)kzg_rd32(c) * XP1), 23) * XP2 + XP3; c += 4; }
        while (c < e) { h = kzg_rotl64(h ^ ((u64)(*c) * XP5), 11) * XP1; c++; }
        h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
        const u32 stored = kzg_rd32(B.src + B.src_off[f] + fr->checksum_off);
        if (stored != (u32)h) status = KZG_E_CHECKSUM_WRONG;
        B.frames[f].xxh_lo = (u32)h;
    }
    B.status[f] = status;
    B.out_size[f] = status == KZG_OK ? len : 0;
}

// ------------------------------------------------------------------------------------------------
static bool g_attr_done = false;

const char* const kzg_stage_names[KZG_N_STAGES] = { "scan", "parse", "link", "huffman", "fse", "offsets", "execute", "finish" };

int kzg_launch_batch(const KzgBatch& b, int sm_count, cudaStream_t stream, cudaEvent_t* ev) {
    const size_t huf_smem = KZG_HUF_SMEM;
    const size_t fse_smem = KZG_FSE_SMEM;
    const size_t exec_smem = b.tile_bytes;
    if (!g_attr_done) {
        cudaFuncSetAttribute(kzg_huf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)huf_smem);
        cudaFuncSetAttribute(kzg_fse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fse_smem);
        cudaFuncSetAttribute(kzg_exec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64);
        g_attr_done = true;
    }
    const u32 n = b.n_frames;
    if (n == 0) return 0;
    int k = 0;
#define KZG_MARK() do { if (ev) cudaEventRecord(ev[k], stream); k++; } while (0)
    const u32 fgrid = (n + 127) / 128;
    KZG_MARK();
    kzg_scan_kernel<<<fgrid, 128, 0, stream>>>(b);
    KZG_MARK();
    const u32 pgrid = (u32)min((u64)(b.cblock_cap + 63) / 64, (u64)sm_count * 16);
    kzg_parse_kernel<<<pgrid ? pgrid : 1, 64, 0, stream>>>(b);
    KZG_MARK();
    kzg_link_kernel<<<fgrid, 128, 0, stream>>>(b);
    KZG_MARK();
    kzg_huf_kernel<<<sm_count * 3, KZG_HUF_WARPS * 32, huf_smem, stream>>>(b);
    KZG_MARK();
    kzg_fse_kernel<<<sm_count, KZG_FSE_WARPS * 32, fse_smem, stream>>>(b);
    KZG_MARK();
    kzg_offsets_kernel<<<fgrid, 128, 0, stream>>>(b);
    KZG_MARK();
    const int exec_per_sm = (int)((227 * 1024) / (exec_smem + 1024));
    kzg_exec_kernel<<<sm_count * (exec_per_sm < 1 ? 1 : exec_per_sm), KZG_EXEC_THREADS, exec_smem, stream>>>(b);
    KZG_MARK();
    kzg_finish_kernel<<<(n * 4 + 127) / 128, 128, 0, stream>>>(b);
    KZG_MARK();
#undef KZG_MARK
    return KZG_N_STAGES;
}
