// B200 (sm_100a) kernels of the batched Zstandard decoder.
//
// Pipeline (one stream, no host round trip until the per-frame status is read back):
//   scan      thread / frame        frame header, block-header walk -> block work list       (stage s1)
//   parse     thread / block        literals + sequences section headers, Huffman weights,
//                                   FSE normalised counts, pool allocation                   (s3/s4 headers)
//   link      thread / frame        treeless / repeat-mode table owners inside a frame
//   fsetab    warp / block          the three FSE decode tables of every block with sequences        (s4 tables)
//   huffman   thread / stream       table build (4 lanes per table) + 4-stream literal decode (s3)
//   fse       thread / block        table build + sequential ZSTD_decodeSequence state machine (s4)
//   offsets   thread / frame        per-block output offsets, repeat-offset hand-over, size checks
//   bulk      CTA / block           raw + RLE blocks, literal-only blocks, long literal runs            (s2, s5)
//   execute   warp / frame          literal + match copy through a shared-memory output window,
//                                   16-byte vectorised flush                                       (s5)
//   finish    4 lanes / frame       XXH64 of the output, checksum compare, status + size      (s6)
//
// The format logic follows RFC 8878 / SURVEY.md Appendix A; it replaces what libzstd does underneath
// the reference call ZSTD_decompressStream (kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:178).
#include "kzg_kernels.cuh"
#include "kzg_stages.cuh"
#include "kzg_huf.cuh"
#include "kzg_fsetab.cuh"
#include "kzg_fse.cuh"
#include "kzg_bulk.cuh"
#include "kzg_exec.cuh"
#include "kzg_exec2.cuh"
#include "kzg_exec3.cuh"
#include "kzg_exec4.cuh"
#include <stdlib.h>
#include <mutex>

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
// huffman: see kzg_huf.cuh (one thread per stream, cp.async bitstream ring, 16 symbols per store)
// ------------------------------------------------------------------------------------------------
// fse: see kzg_fse.cuh (one thread per block, lane-interleaved tables, cp.async bitstream ring)
__global__ void __launch_bounds__(128) kzg_offsets_kernel(KzgBatch B) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.n_frames || (B.ctr->overflow & 3u)) return;
    kzg_stage_offsets(B, i);
}

// ------------------------------------------------------------------------------------------------
// execute: see kzg_exec.cuh (one warp per frame, shared-memory output window, vector flushes)

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
        // literals libzstd would have made up (kzg_huf_verdict): its checksum cannot match; without one the frame is rejected
        if (status == KZG_OK && fr->lit_garbage) status = fr->checksum_off != KZG_NONE ? (u32)KZG_E_CHECKSUM_WRONG : (u32)KZG_E_CORRUPTION;
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
        // 16 independent loads in flight per lane: the rounds themselves are a serial chain, the loads are not
        if (((uintptr_t)s & 7) == 0) {
            const u64* s64 = reinterpret_cast<const u64*>(s);
            for (; k + 16 <= stripes; k += 16) {
                u64 x[16];
#pragma unroll
                for (int j = 0; j < 16; j++) x[j] = __ldcs(s64 + 4 * (k + j));
#pragma unroll
                for (int j = 0; j < 16; j++) acc = kzg_xxround(acc, x[j]);
            }
        } else {
            for (; k + 8 <= stripes; k += 8) {
                u64 x[8];
#pragma unroll
                for (int j = 0; j < 8; j++) x[j] = kzg_ld64_any(s + 32 * (k + j));
#pragma unroll
                for (int j = 0; j < 8; j++) acc = kzg_xxround(acc, x[j]);
            }
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
    // on an error: the index of the block it belongs to (the blocks before it are good; the streaming shim hands them over
    // first, as libzstd does) -- except for a checksum that does not match: the whole content is there
    // (bit 63 marks the one checksum verdict that comes without the content: literals libzstd would have made up)
    // (bit 62: the error is that kind, whatever the code -- a segment has no checksum of its own, its frame may have one)
    B.out_size[f] = status == KZG_OK || (status == KZG_E_CHECKSUM_WRONG && !fr->lit_garbage) ? len
                    : ((u64)fr->fail_block | (status == KZG_E_CHECKSUM_WRONG ? 1ull << 63 : 0ull) | (fr->lit_garbage ? 1ull << 62 : 0ull));
}

// ------------------------------------------------------------------------------------------------
// XXH64 over a stream that arrives in segments: the four accumulators on four lanes as above; up to 31 bytes that do
// not fill a 32-byte stripe wait in the state.  The digest (as if the stream ended here) is recomputed every time.
__global__ void __launch_bounds__(32) kzg_xxh_segment_kernel(KzgXxhState* S, const u8* data, u64 len) {
    const u32 q = threadIdx.x & 31;
    const u32 t = (u32)(S->total & 31ull);                 // bytes waiting in the state
    const u64 all = (u64)t + len, stripes = all >> 5;
    u64 acc = q < 4 ? S->acc[q] : 0ull;
    auto logical = [&](u64 p) -> u32 { return p < t ? (u32)S->tail[p] : (u32)data[p - t]; };
    if (q < 4 && stripes) {
        u64 k = 0;
        if (t) {                                           // the first stripe starts in the waiting bytes
            u64 x = 0;
            for (int b = 0; b < 8; b++) x |= (u64)logical(8 * q + b) << (8 * b);
            acc = kzg_xxround(acc, x);
            k = 1;
        }
        const u8* s = data + (t ? 32 - t : 0) + 8 * q;      // stripe k (k >= 1 if t) starts at data offset 32 * k - t
        const u64 rest = stripes - k;
        u64 j = 0;
        for (; j + 16 <= rest; j += 16) {
            u64 x[16];
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = kzg_ld64_any(s + 32 * (j + i));
#pragma unroll
            for (int i = 0; i < 16; i++) acc = kzg_xxround(acc, x[i]);
        }
        for (; j < rest; j++) acc = kzg_xxround(acc, kzg_ld64_any(s + 32 * j));
    }
    __syncwarp();
    // the bytes that wait now: logical positions [32 * stripes, all)
    const u32 nt = (u32)(all & 31ull);
    u32 tb = 0;
    if (q < nt) tb = logical((stripes << 5) + q);
    __syncwarp();
    if (q < nt) S->tail[q] = (u8)tb;
    if (q < 4) S->acc[q] = acc;
    const u64 total = S->total + len;
    const u64 v1 = __shfl_sync(0xFFFFFFFFu, acc, 0), v2 = __shfl_sync(0xFFFFFFFFu, acc, 1), v3 = __shfl_sync(0xFFFFFFFFu, acc, 2),
              v4 = __shfl_sync(0xFFFFFFFFu, acc, 3);
    __syncwarp();
    if (q == 0) {
        S->total = total;
        u64 h;
        if (total >= 32) {
            h = kzg_rotl64(v1, 1) + kzg_rotl64(v2, 7) + kzg_rotl64(v3, 12) + kzg_rotl64(v4, 18);
            h = (h ^ kzg_xxround(0, v1)) * XP1 + XP4;
            h = (h ^ kzg_xxround(0, v2)) * XP1 + XP4;
            h = (h ^ kzg_xxround(0, v3)) * XP1 + XP4;
            h = (h ^ kzg_xxround(0, v4)) * XP1 + XP4;
        } else h = XP5;
        h += total;
        const u8* c = S->tail;
        const u8* e = c + nt;
        while (e - c >= 8) { h = kzg_rotl64(h ^ kzg_xxround(0, kzg_rd64(c)), 27) * XP1 + XP4; c += 8; }
        if (e - c >= 4) { h = kzg_rotl64(h ^ ((u64)kzg_rd32(c) * XP1), 23) * XP2 + XP3; c += 4; }
        while (c < e) { h = kzg_rotl64(h ^ ((u64)(*c) * XP5), 11) * XP1; c++; }
        h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
        S->digest = (u32)h;
    }
}

void kzg_launch_xxh_segment(KzgXxhState* d_state, const u8* d_data, u64 len, cudaStream_t stream) {
    kzg_xxh_segment_kernel<<<1, 32, 0, stream>>>(d_state, d_data, len);
}

// ------------------------------------------------------------------------------------------------
const char* const kzg_stage_names[KZG_N_STAGES] = { "scan", "parse", "link", "fsetab", "fse", "huffman", "offsets", "bulk", "execute", "execute_runs", "finish" };

// Which execute kernel runs first: the first design's hot variant (kzg_exec.cuh, lane per sequence), or with KZG_EXEC=2
// the output-centric second design (kzg_exec2.cuh) and its ring size KZG_EXEC_TILE.  The second design is bit-exact and
// tested but slower on every corpus measured (profiles/r02_exec2_experiment.txt), so it is not the default.  Read at
// every launch so that one process can compare them.
struct KzgExecChoice { int design; u32 tile; };
static KzgExecChoice kzg_exec_choice() {
    KzgExecChoice r{ 1, 0u };
    if (const char* e = getenv("KZG_EXEC")) { const int d = atoi(e); r.design = d >= 1 && d <= 4 ? d : 1; }
    if (const char* e = getenv("KZG_EXEC_TILE")) r.tile = (u32)strtoul(e, nullptr, 10);
    if (r.design == 2 && r.tile != 4096u) r.tile = 8192u;
    if (r.design == 3 && r.tile != 512u && r.tile != 2048u) r.tile = 1024u;
    return r;
}

// Dynamic shared-memory opt-in is per device (per primary context): done once for every device a context is created on.
int kzg_kernels_init_device(int device) {
    static std::mutex mu;
    static u64 done_mask[4] = { 0, 0, 0, 0 };
    std::lock_guard<std::mutex> lock(mu);
    if (device >= 0 && device < 256 && (done_mask[device >> 6] >> (device & 63) & 1ull)) return 0;
    cudaError_t e = cudaSuccess;
    auto opt = [&](const void* fn, size_t bytes) { if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); };
    opt((const void*)kzg_huf_kernel<KZG_HUF_L1_BITS, KZG_HUF_L2CAP, false>, KZG_HUF_SMEM);
    opt((const void*)kzg_huf_kernel<KZG_HUF_MAXBITS, 0, true>, KZG_HUF_SMEM_FULL);
    opt((const void*)kzg_fse_kernel, KZG_FSE_SMEM_TOTAL);
    opt((const void*)kzg_exec_kernel<false>, KZG_EX_SMEM);
    opt((const void*)kzg_exec_kernel<true>, KZG_EX_SMEM);
    opt((const void*)kzg_exec2_kernel<4096>, KZG_X2_WARPS * KzgX2<4096>::BYTES);
    opt((const void*)kzg_exec2_kernel<8192>, KZG_X2_WARPS * KzgX2<8192>::BYTES);
    opt((const void*)kzg_exec4_kernel<false>, KZG_X4_SMEM);
    opt((const void*)kzg_exec4_kernel<true>, KZG_X4_SMEM);
    opt((const void*)kzg_exec3_kernel<512>, KzgX3<512>::BYTES);
    opt((const void*)kzg_exec3_kernel<1024>, KzgX3<1024>::BYTES);
    opt((const void*)kzg_exec3_kernel<2048>, KzgX3<2048>::BYTES);
    if (e != cudaSuccess) { cudaGetLastError(); return -1; }
    if (device >= 0 && device < 256) done_mask[device >> 6] |= 1ull << (device & 63);
    return 0;
}

int kzg_launch_batch(const KzgBatch& b_in, int sm_count, cudaStream_t stream, cudaEvent_t* ev) {
    KzgBatch b = b_in;
    const size_t huf_smem = KZG_HUF_SMEM;
    // literal warps riding along in the sequence kernel (KZG_FSE_LITWARPS, 0..3: development knob, read at every launch)
    u32 nlit = KZG_FSE_LIT_WARPS_MAX;
    if (const char* e = getenv("KZG_FSE_LITWARPS")) { const long v = strtol(e, nullptr, 10); if (v >= 0 && v <= KZG_FSE_LIT_WARPS_MAX) nlit = (u32)v; }
    const size_t fse_smem = KZG_FSE_SMEM_TOTAL_OF(nlit);
    const size_t exec_smem = KZG_EX_SMEM;
    const u32 n = b.n_frames;
    if (n == 0) return 0;
    // Long literal streams decode faster with full single-level tables (shorter chain per symbol), but only 48 blocks fit
    // an SM then; with the compact tables 128 do.  Few blocks: the chain counts.  Many: the number of waves does.
    b.huf_long_min = n > (u32)sm_count * 48u ? 0xFFFFFFFFu : KZG_HUF_LONG_STREAMS;
    if (const char* e = getenv("KZG_HUF_LONG_MIN")) b.huf_long_min = (u32)strtoul(e, nullptr, 10);
    b.lit_ride = 0;
    if (const char* e = getenv("KZG_FSE_RIDE")) b.lit_ride = (u32)strtoul(e, nullptr, 10);     // development knob
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
#ifndef KZG_TAB_GRID_PER_SM
#define KZG_TAB_GRID_PER_SM 8
#endif
    kzg_fsetab_kernel<<<sm_count * KZG_TAB_GRID_PER_SM, KZG_TAB_WARPS * 32, 0, stream>>>(b);
    KZG_MARK();
    kzg_fse_kernel<<<sm_count, KZG_FSE_THREADS_OF(nlit), fse_smem, stream>>>(b);          // sequences, and literals on the side
    KZG_MARK();
    {   // the literals that are left: 16 warps per SM with the compact two-level tables, then the blocks that launch
        // leaves (long streams, oversized second levels) with full tables, 6 warps per SM
        const int per_sm = (int)((228 * 1024) / (huf_smem + 1024));
        kzg_huf_kernel<KZG_HUF_L1_BITS, KZG_HUF_L2CAP, false><<<sm_count * per_sm, KZG_HUF_WARPS * 32, huf_smem, stream>>>(b);
        kzg_huf_kernel<KZG_HUF_MAXBITS, 0, true><<<sm_count * 3, KZG_HUF_WARPS * 32, KZG_HUF_SMEM_FULL, stream>>>(b);
    }
    KZG_MARK();
    kzg_offsets_kernel<<<fgrid, 128, 0, stream>>>(b);
    KZG_MARK();
    kzg_bulk_kernel<<<sm_count * 4, KZG_BULK_THREADS, 0, stream>>>(b);
    KZG_MARK();
    int exec_per_sm = (int)((227 * 1024) / (exec_smem + 1024));
    if (exec_per_sm > KZG_EX_MINBLOCKS) exec_per_sm = KZG_EX_MINBLOCKS;
    KzgExecChoice ch = kzg_exec_choice();
    if (ch.design == 3 && b.dict_size) ch.design = 1;          // history before the frame: the first design reads it
    // Few frames cannot fill the GPU one warp each: below this many, every frame gets a CTA (kzg_exec4.cuh).
    u32 cta_all_below = (u32)sm_count * 2u;
    if (const char* e = getenv("KZG_EXEC_CTA_BELOW")) cta_all_below = (u32)strtoul(e, nullptr, 10);
    if (ch.design == 1 && n < cta_all_below) ch.design = 4;
    bool big_by_cta = true;                                     // deferred frames: the CTA kernel (0: the first design's full variant)
    if (const char* e = getenv("KZG_EXEC_BIG")) big_by_cta = atoi(e) != 0;
    b.big_by_cta = big_by_cta ? 1u : 0u;
    if (ch.design == 4) kzg_exec4_kernel<true><<<min((u32)sm_count, n), KZG_X4_THREADS, KZG_X4_SMEM, stream>>>(b);
    else if (ch.design == 1) kzg_exec_kernel<false><<<sm_count * exec_per_sm, KZG_EX_WARPS * 32, exec_smem, stream>>>(b);
    else if (ch.design == 3) {
        // a thread per frame: one warp per CTA, as many CTAs per SM as the rings allow
#define KZG_X3_LAUNCH(T) do { const size_t sm_ = KzgX3<T>::BYTES; const int per_sm_ = (int)((228 * 1024) / (sm_ + 1024)); \
        kzg_exec3_kernel<T><<<sm_count * per_sm_, 32, sm_, stream>>>(b); } while (0)
        if (ch.tile == 512u) KZG_X3_LAUNCH(512);
        else if (ch.tile == 2048u) KZG_X3_LAUNCH(2048);
        else KZG_X3_LAUNCH(1024);
#undef KZG_X3_LAUNCH
    } else {
        // one warp per frame, KZG_X2_WARPS per CTA, as many CTAs per SM as the ring size allows
#define KZG_X2_LAUNCH(T) do { const size_t sm_ = KZG_X2_WARPS * KzgX2<T>::BYTES; const int per_sm_ = (int)((228 * 1024) / (sm_ + 1024)); \
        kzg_exec2_kernel<T><<<sm_count * (per_sm_ > 16 ? 16 : per_sm_), KZG_X2_WARPS * 32, sm_, stream>>>(b); } while (0)
        // (rings of 16 and 32 KiB were measured too, profiles/r02_exec2_experiment.txt: slower still, and the 32 KiB build
        // failed one GPU parity test that the emulator passes -- not pursued, the instantiations are gone)
        if (ch.tile == 4096u) KZG_X2_LAUNCH(4096);
        else KZG_X2_LAUNCH(8192);
#undef KZG_X2_LAUNCH
    }
    KZG_MARK();
    // deferred frames only: long runs -> the first design's full variant; many blocks -> a CTA per frame
    int launched = KZG_N_STAGES + 1;                            // the Huffman stage is two launches
    if (ch.design == 4) launched--;                             // nothing was deferred: one launch fewer
    else {
        kzg_exec_kernel<true><<<sm_count * exec_per_sm, KZG_EX_WARPS * 32, exec_smem, stream>>>(b);
        if (big_by_cta) { kzg_exec4_kernel<false><<<sm_count, KZG_X4_THREADS, KZG_X4_SMEM, stream>>>(b); launched++; }
    }
    KZG_MARK();
    kzg_finish_kernel<<<(n * 4 + 127) / 128, 128, 0, stream>>>(b);
    KZG_MARK();
#undef KZG_MARK
    return launched;                // kernels launched
}
