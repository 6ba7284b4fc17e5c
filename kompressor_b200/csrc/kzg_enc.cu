// C ABI of the batch frame writer (include/kzg.h, "Layer A, compress side"): kzg_enc_* and kzg_compress_batch_*.
// Kernels and format notes: kzg_enc.cuh.
#include "../../include/kzg.h"
#include "kzg_enc.cuh"

#include <cuda_runtime.h>
#include <mutex>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define KZG_ENC_MAX_SUB 32                    // host-pointer API: sub-batches pipelined over the copy engines
#define KZG_ENC_SUB_BYTES (128ull << 20)      // input bytes per sub-batch

namespace {
struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + (bytes >> 3) + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return -1; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
}  // namespace

struct kzg_enc {
    int device = 0;
    int sm_count = 0;
    std::mutex mu;
    cudaStream_t stream = nullptr, in_stream = nullptr, out_stream = nullptr;     // kernels; H2D; D2H of the host-pointer API
    cudaEvent_t ev_in[KZG_ENC_MAX_SUB], ev_cmp[KZG_ENC_MAX_SUB];
    Buf tables, seq, chunks, tickets, lit;       // workspace
    Buf st_src, st_dst, st_tab;                  // staging of the host-pointer API
    cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
    float stage_ms[4] = { 0, 0, 0, 0 };
    bool profiling = false;
};

#define CKE(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "kzg: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -(int)KZG_E_MEMORY; } } while (0)

extern "C" size_t kzg_compress_bound(size_t src_size) { return (size_t)kzg_enc_bound(src_size); }

extern "C" int kzg_enc_create(kzg_enc** out, int device) {
    if (!out) return -(int)KZG_E_GENERIC;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return -(int)KZG_E_MEMORY; }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= count) return -(int)KZG_E_GENERIC;
    CKE(cudaSetDevice(device));
    kzg_enc* e = new (std::nothrow) kzg_enc();
    if (!e) return -(int)KZG_E_MEMORY;
    e->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete e; return -(int)KZG_E_MEMORY; }
    e->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->in_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->out_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); delete e; return -(int)KZG_E_MEMORY; }
    for (int k = 0; k < 5; k++) cudaEventCreate(&e->ev[k]);
    for (int k = 0; k < KZG_ENC_MAX_SUB; k++) {
        cudaEventCreateWithFlags(&e->ev_in[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&e->ev_cmp[k], cudaEventDisableTiming);
    }
    KzgEncTables host_tables;
    kzg_enc_build_tables(host_tables);
    if (e->tables.ensure(sizeof(KzgEncTables)) || e->tickets.ensure(64) ||
        cudaMemcpy(e->tables.p, &host_tables, sizeof host_tables, cudaMemcpyHostToDevice) != cudaSuccess) {
        kzg_enc_destroy(e);
        return -(int)KZG_E_MEMORY;
    }
    cudaFuncSetAttribute(kzg_enc_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kzg_enc_match_smem(KZG_BLOCK_MAX));
    *out = e;
    return 0;
}

extern "C" void kzg_enc_destroy(kzg_enc* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) { cudaStreamSynchronize(e->stream); cudaStreamDestroy(e->stream); }
    if (e->in_stream) cudaStreamDestroy(e->in_stream);
    if (e->out_stream) cudaStreamDestroy(e->out_stream);
    for (int k = 0; k < 5; k++) if (e->ev[k]) cudaEventDestroy(e->ev[k]);
    for (int k = 0; k < KZG_ENC_MAX_SUB; k++) { if (e->ev_in[k]) cudaEventDestroy(e->ev_in[k]); if (e->ev_cmp[k]) cudaEventDestroy(e->ev_cmp[k]); }
    e->tables.release(); e->seq.release(); e->chunks.release(); e->tickets.release(); e->lit.release();
    e->st_src.release(); e->st_dst.release(); e->st_tab.release();
    delete e;
}

extern "C" int kzg_enc_set_profiling(kzg_enc* e, int enable) { if (!e) return -(int)KZG_E_GENERIC; e->profiling = enable != 0; return 0; }

extern "C" int kzg_enc_last_stage_ms(kzg_enc* e, float* ms, const char** names, int max_stages) {
    static const char* const kNames[4] = { "match", "literals", "encode", "finish" };
    if (!e) return -(int)KZG_E_GENERIC;
    int n = max_stages < 4 ? max_stages : 4;
    for (int k = 0; k < n; k++) { if (ms) ms[k] = e->stage_ms[k]; if (names) names[k] = kNames[k]; }
    return n;
}

// Enqueues the three kernels on `s`. All pointers are device pointers.
static int enqueue(kzg_enc* e, const void* d_src, const uint64_t* d_src_off, const uint64_t* d_src_size, void* d_dst,
                   const uint64_t* d_dst_off, const uint64_t* d_dst_cap, uint64_t n, uint64_t max_chunk, int flags,
                   uint32_t* d_status, uint64_t* d_out_size, cudaStream_t s) {
    if (n == 0) return 0;
    if (n > 0xFFFFFFF0ull) return -(int)KZG_E_GENERIC;
    if (max_chunk > KZG_BLOCK_MAX) max_chunk = KZG_BLOCK_MAX;       // larger chunks are reported per chunk (srcSize_wrong)
    const u32 seq_cap = KZG_ENC_WARPS * (u32)(max_chunk / (6 * KZG_ENC_WARPS) + 32);     // per part: one sequence per 6 bytes at most
    const u32 lit_cap = (u32)((max_chunk + 15) & ~15ull) + 16;
    if (e->seq.ensure((size_t)n * seq_cap * 8) || e->chunks.ensure((size_t)n * sizeof(KzgEncChunk)) || e->lit.ensure((size_t)n * lit_cap))
        return -(int)KZG_E_MEMORY;
    KzgEncBatch B;
    memset(&B, 0, sizeof B);
    B.src = (const u8*)d_src; B.src_off = d_src_off; B.src_size = d_src_size;
    B.dst = (u8*)d_dst; B.dst_off = d_dst_off; B.dst_cap = d_dst_cap;
    B.status = d_status; B.out_size = d_out_size; B.n = (u32)n; B.flags = (u32)flags;
    B.seq = (u64*)e->seq.p; B.seq_cap = seq_cap; B.chunks = (KzgEncChunk*)e->chunks.p;
    B.tab = (const KzgEncTables*)e->tables.p; B.tickets = (u32*)e->tickets.p;
    B.lit = (u8*)e->lit.p; B.lit_cap = lit_cap;
    CKE(cudaMemsetAsync(e->tickets.p, 0, 64, s));
    // the chunk area of the match kernel's shared memory: 64 KiB (two chunks per SM) unless the batch has larger chunks
    const u32 chunk_cap = max_chunk <= 65536 ? 65536u : KZG_BLOCK_MAX;
    const int smem = (int)kzg_enc_match_smem(chunk_cap);
    const int per_sm = (227 * 1024) / (smem + 1024);
    const bool prof = e->profiling;
    if (prof) cudaEventRecord(e->ev[0], s);
    kzg_enc_match_kernel<<<e->sm_count * per_sm, KZG_ENC_WARPS * 32, smem, s>>>(B, chunk_cap);
    if (prof) cudaEventRecord(e->ev[1], s);
    kzg_enc_literals_kernel<<<e->sm_count * 8, KZG_ENC_HUF_THREADS, 0, s>>>(B);
    if (prof) cudaEventRecord(e->ev[2], s);
    kzg_enc_encode_kernel<<<e->sm_count * 8, KZG_ENC_EWARPS * 32, 0, s>>>(B);
    if (prof) cudaEventRecord(e->ev[3], s);
    kzg_enc_finish_kernel<<<e->sm_count * 8, 128, 0, s>>>(B);
    if (prof) cudaEventRecord(e->ev[4], s);
    CKE(cudaGetLastError());
    return 4;
}

extern "C" int kzg_compress_batch_device(kzg_enc* e, const void* d_src, const uint64_t* d_src_off, const uint64_t* d_src_size,
                                         void* d_dst, const uint64_t* d_dst_off, const uint64_t* d_dst_cap, uint64_t n,
                                         uint64_t max_chunk, int flags, uint32_t* d_status, uint64_t* d_out_size, void* stream) {
    if (!e) return -(int)KZG_E_GENERIC;
    std::lock_guard<std::mutex> lock(e->mu);
    CKE(cudaSetDevice(e->device));
    return enqueue(e, d_src, d_src_off, d_src_size, d_dst, d_dst_off, d_dst_cap, n, max_chunk, flags, d_status, d_out_size,
                   stream ? (cudaStream_t)stream : e->stream);
}

extern "C" int kzg_enc_finish(kzg_enc* e, void* stream) {
    if (!e) return -(int)KZG_E_GENERIC;
    std::lock_guard<std::mutex> lock(e->mu);
    CKE(cudaSetDevice(e->device));
    CKE(cudaStreamSynchronize(stream ? (cudaStream_t)stream : e->stream));
    if (e->profiling)
        for (int k = 0; k < 4; k++) if (cudaEventElapsedTime(&e->stage_ms[k], e->ev[k], e->ev[k + 1]) != cudaSuccess) { cudaGetLastError(); e->stage_ms[k] = 0; }
    return 0;
}

extern "C" int kzg_compress_batch_host(kzg_enc* e, const void* src, const uint64_t* src_off, const uint64_t* src_size,
                                       void* dst, const uint64_t* dst_off, const uint64_t* dst_cap, uint64_t n, int flags,
                                       uint32_t* status, uint64_t* out_size) {
    if (!e) return -(int)KZG_E_GENERIC;
    if (n == 0) return 0;
    std::unique_lock<std::mutex> lock(e->mu);
    CKE(cudaSetDevice(e->device));
    u64 src_lo = ~0ull, src_hi = 0, dst_lo = ~0ull, dst_hi = 0, max_chunk = 0;
    for (u64 i = 0; i < n; i++) {
        if (src_off[i] < src_lo) src_lo = src_off[i];
        if (src_off[i] + src_size[i] > src_hi) src_hi = src_off[i] + src_size[i];
        if (dst_off[i] < dst_lo) dst_lo = dst_off[i];
        if (dst_off[i] + dst_cap[i] > dst_hi) dst_hi = dst_off[i] + dst_cap[i];
        if (src_size[i] > max_chunk) max_chunk = src_size[i];
    }
    const u64 src_span = src_hi - src_lo, dst_span = dst_hi - dst_lo;
    const u64 tab_bytes = n * (5 * 8 + 4);
    if (e->st_src.ensure(src_span + 64) || e->st_dst.ensure(dst_span + 64) || e->st_tab.ensure(tab_bytes + 64)) return -(int)KZG_E_MEMORY;
    u64* t_off = (u64*)e->st_tab.p;
    u64* t_size = t_off + n;
    u64* t_doff = t_size + n;
    u64* t_dcap = t_doff + n;
    u64* t_osize = t_dcap + n;
    u32* t_status = (u32*)(t_osize + n);
    u64* h_tab = (u64*)malloc(n * 4 * 8);
    if (!h_tab) return -(int)KZG_E_MEMORY;
    for (u64 i = 0; i < n; i++) {
        h_tab[i] = src_off[i] - src_lo; h_tab[n + i] = src_size[i];
        h_tab[2 * n + i] = dst_off[i] - dst_lo; h_tab[3 * n + i] = dst_cap[i];
    }
    // Sub-batches of consecutive chunks, pipelined over three streams: while sub-batch k is compressed, the input of
    // k+1 is on its way in and the frames of k-1 are on their way out.  The kernels of successive sub-batches run
    // back to back on one stream and share the workspace.
    u64 bounds[KZG_ENC_MAX_SUB + 1];
    u64 nsub = 0;
    bounds[0] = 0;
    {
        u64 total_src = 0;
        for (u64 i = 0; i < n; i++) total_src += src_size[i];
        u64 goal = KZG_ENC_SUB_BYTES;
        while (total_src / goal >= KZG_ENC_MAX_SUB) goal <<= 1;
        u64 acc = 0;
        for (u64 i = 0; i < n; i++) {
            acc += src_size[i];
            if (acc >= goal || i + 1 == n) { bounds[++nsub] = i + 1; acc = 0; }
        }
    }
    cudaStream_t s = e->stream, s_in = e->in_stream, s_out = e->out_stream;
    bool ok = cudaMemcpyAsync(e->st_tab.p, h_tab, n * 4 * 8, cudaMemcpyHostToDevice, s_in) == cudaSuccess;
    int rc = 0;
    for (u64 k = 0; ok && k < nsub; k++) {
        const u64 f0 = bounds[k], f1 = bounds[k + 1];
        u64 lo = ~0ull, hi = 0, mc = 0;
        for (u64 i = f0; i < f1; i++) {
            if (src_off[i] < lo) lo = src_off[i];
            if (src_off[i] + src_size[i] > hi) hi = src_off[i] + src_size[i];
            if (src_size[i] > mc) mc = src_size[i];
        }
        ok = (hi <= lo || cudaMemcpyAsync((u8*)e->st_src.p + (lo - src_lo), (const u8*)src + lo, hi - lo, cudaMemcpyHostToDevice, s_in) == cudaSuccess) &&
             cudaEventRecord(e->ev_in[k], s_in) == cudaSuccess && cudaStreamWaitEvent(s, e->ev_in[k], 0) == cudaSuccess;
        if (!ok) break;
        rc = enqueue(e, e->st_src.p, t_off + f0, t_size + f0, e->st_dst.p, t_doff + f0, t_dcap + f0, f1 - f0, max_chunk, flags,
                     t_status + f0, t_osize + f0, s);
        if (rc < 0) break;
        ok = cudaEventRecord(e->ev_cmp[k], s) == cudaSuccess && cudaStreamWaitEvent(s_out, e->ev_cmp[k], 0) == cudaSuccess;
        if (nsub > 1) {
            // frame sizes are not known on the host yet: whole capacity regions, adjacent ones coalesced
            u64 run_lo = dst_off[f0], run_hi = dst_off[f0] + dst_cap[f0];
            for (u64 i = f0 + 1; ok && i <= f1; i++) {
                if (i < f1 && dst_off[i] == run_hi) { run_hi += dst_cap[i]; continue; }
                if (run_hi > run_lo)
                    ok = cudaMemcpyAsync((u8*)dst + run_lo, (const u8*)e->st_dst.p + (run_lo - dst_lo), run_hi - run_lo, cudaMemcpyDeviceToHost, s_out) == cudaSuccess;
                if (i < f1) { run_lo = dst_off[i]; run_hi = dst_off[i] + dst_cap[i]; }
            }
        }
    }
    // status and sizes once, at the end: the caller's arrays may be pageable, and a copy into pageable memory makes
    // the host wait for everything enqueued before it
    if (ok && rc >= 0)
        ok = cudaMemcpyAsync(status, t_status, n * 4, cudaMemcpyDeviceToHost, s_out) == cudaSuccess &&
             cudaMemcpyAsync(out_size, t_osize, n * 8, cudaMemcpyDeviceToHost, s_out) == cudaSuccess;
    const bool synced = cudaStreamSynchronize(s_in) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess && cudaStreamSynchronize(s_out) == cudaSuccess;
    ok = ok && synced;
    if (ok && rc >= 0 && nsub == 1) {
        // one sub-batch: copy back either what each frame occupies (sizes are known now) or, when there are many frames,
        // whole capacity regions with adjacent ones coalesced (one copy per run instead of one per frame)
        u64 total_out = 0, total_cap = 0;
        for (u64 i = 0; i < n; i++) { total_out += status[i] == 0 ? out_size[i] : 0; total_cap += dst_cap[i]; }
        const bool exact = (double)n * 3e-6 + (double)total_out / 50e9 < (double)total_cap / 50e9;
        u64 run_lo = 0, run_hi = 0;
        bool open = false;
        for (u64 i = 0; ok && i <= n; i++) {
            const u64 len = i < n ? (exact ? (status[i] == 0 ? out_size[i] : 0) : dst_cap[i]) : 0;
            const bool have = i < n && len > 0;
            if (have && open && dst_off[i] == run_hi) { run_hi += len; continue; }
            if (open) ok = cudaMemcpyAsync((u8*)dst + run_lo, (const u8*)e->st_dst.p + (run_lo - dst_lo), run_hi - run_lo, cudaMemcpyDeviceToHost, s_out) == cudaSuccess;
            open = have;
            if (have) { run_lo = dst_off[i]; run_hi = dst_off[i] + len; }
        }
        ok = ok && cudaStreamSynchronize(s_out) == cudaSuccess;
    }
    free(h_tab);
    if (rc < 0) return rc;
    if (!ok) { cudaGetLastError(); return -(int)KZG_E_MEMORY; }
    if (e->profiling)
        for (int k = 0; k < 4; k++) if (cudaEventElapsedTime(&e->stage_ms[k], e->ev[k], e->ev[k + 1]) != cudaSuccess) { cudaGetLastError(); e->stage_ms[k] = 0; }
    return 0;
}
