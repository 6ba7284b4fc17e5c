// C ABI, layer B: streaming shim with libzstd semantics (include/kzg.h).
//
// Replaces, one for one, the six libzstd functions the reference binds for decompression
// (kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:58-73,123-196 and
//  .../nativeMain/kotlin/com/ensody/kompressor/zstd/ZstdDecompressor.native.kt:29-76).
//
// A kzg_DCtx accumulates input until at least one complete frame is buffered, decodes every complete
// frame it holds as ONE GPU batch (in transform(ByteArray) the first call exposes the whole array,
// SliceTransform.kt:34,39, so all concatenated frames become a single batch), then drains the output
// over the following calls.  It never withholds a complete frame: the reference never signals end of
// input to the decompressor (ZstdDecompressor.jvm.kt:26-36 drops `finish`).
//
// Behind pipe() (SliceTransformRawSource.kt:32-55, 8 KiB slices) one handle completes at most one frame per
// call, and one frame is a latency-bound GPU batch (~1.3 ms).  Many handles are active at once in a server
// (one per HTTP body, SURVEY.md section 8(b) "Threading"), so the shim merges across handles: calls that arrive
// while a batch is running are queued and the next batch decodes all of them together (group commit; the
// thread that finds the device idle leads).  A batch of 64 frames takes about as long as a batch of one.
//   * small requests are gathered into pinned staging (one H2D, one D2H) and each caller copies its own
//     frames out afterwards, in parallel; two staging generations alternate so the next batch does not wait
//     for the slowest reader;
//   * large requests (the transform(ByteArray) shape) run alone, straight from the caller's buffer into the
//     handle's pinned output arena; kzg_decompress_batch_host pipelines H2D / decode / D2H over sub-batches;
//   * output arenas and staging come from a process-wide pool of pinned allocations (cudaMallocHost is slow,
//     and fresh pageable memory costs a page fault per 4 KiB).
#include "../../include/kzg.h"
#include "kzg_parse.cuh"

#include <condition_variable>
#include <mutex>
#include <new>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

inline size_t err(u32 code) { return (size_t)0 - (size_t)code; }

// ------------------------------------------------------------------ pinned host memory, pooled
struct Arena { u8* p = nullptr; size_t cap = 0; bool pinned = true; };

class PinnedPool {
    std::mutex mu_;
    std::vector<Arena> free_;
    size_t kept_ = 0;
    static constexpr size_t kKeepBytes = 2ull << 30;      // at most this much idle pinned memory is kept for reuse

    static size_t size_class(size_t bytes) {               // powers of two up to 64 MiB, then multiples of 64 MiB
        size_t c = 64u << 10;
        while (c < bytes && c < (64u << 20)) c <<= 1;
        if (c >= bytes) return c;
        return (bytes + (64u << 20) - 1) & ~(size_t)((64u << 20) - 1);
    }

public:
    Arena acquire(size_t bytes) {
        const size_t want = size_class(bytes ? bytes : 1);
        {
            std::lock_guard<std::mutex> lock(mu_);
            size_t best = free_.size();
            for (size_t i = 0; i < free_.size(); i++)
                if (free_[i].cap >= want && free_[i].cap <= 2 * want && (best == free_.size() || free_[i].cap < free_[best].cap)) best = i;
            if (best != free_.size()) {
                Arena a = free_[best];
                free_[best] = free_.back();
                free_.pop_back();
                kept_ -= a.cap;
                return a;
            }
        }
        Arena a;
        a.p = (u8*)kzg_host_alloc(want);
        if (!a.p) {                                        // make room and try once more
            trim(0);
            a.p = (u8*)kzg_host_alloc(want);
        }
        if (!a.p) {                                        // no pinned memory left: pageable memory still works (slower copies)
            a.p = (u8*)malloc(want);
            a.pinned = false;
        }
        a.cap = a.p ? want : 0;
        return a;
    }
    void release(Arena& a) {
        if (!a.p) return;
        if (!a.pinned) { free(a.p); a = Arena(); return; }
        {
            std::lock_guard<std::mutex> lock(mu_);
            if (kept_ + a.cap <= kKeepBytes) { free_.push_back(a); kept_ += a.cap; a = Arena(); return; }
        }
        kzg_host_free(a.p);
        a = Arena();
    }
    void trim(size_t keep) {
        std::vector<Arena> drop;
        {
            std::lock_guard<std::mutex> lock(mu_);
            while (kept_ > keep && !free_.empty()) { drop.push_back(free_.back()); kept_ -= free_.back().cap; free_.pop_back(); }
        }
        for (Arena& a : drop) kzg_host_free(a.p);
    }
};

PinnedPool g_pool;

// ------------------------------------------------------------------ cross-handle batcher
struct Gen {                       // one generation of pinned staging for merged batches
    Arena src, dst;
    int readers = 0;               // requests that have not copied their output out yet (guarded by Batcher::mu)
    std::vector<uint64_t> off, size, doff, cap, osize;
    std::vector<uint32_t> status;
};

struct Req {
    const u8* src = nullptr;                     // frames at src + off[i], size[i]
    const uint64_t *off = nullptr, *size = nullptr, *content = nullptr, *doff = nullptr;
    size_t n = 0;
    uint64_t src_bytes = 0, dst_bytes = 0;       // sums (dst: 16-byte aligned strides, = doff[n-1] + stride)
    u8* dst = nullptr;                           // output of frame i at dst + doff[i]
    uint32_t* status = nullptr;
    uint64_t* osize = nullptr;
    const std::vector<u8>* dict = nullptr;
    int rc = 0;
    bool taken = false, done = false;             // taken: part of a running batch
    Gen* gen = nullptr;                          // merged run: the output waits in gen->dst at gen_base + doff[i]
    uint64_t gen_base = 0;
};

// A lane is one device context (streams + workspace) with its two staging generations.  Several lanes let
// independent batches overlap on the GPU: a batch of a few frames occupies a few SMs for ~1 ms, so with two to
// four active handles each call gets its own lane, and merging sets in once all lanes are busy.
struct Lane {
    kzg_ctx* ctx = nullptr;
    bool busy = false;
    Gen gens[2];
    unsigned flip = 0;
};

class Batcher {
    static constexpr int kLanes = 4;
    std::mutex mu_;
    std::condition_variable cv_;
    Lane lanes_[kLanes];
    std::vector<Req*> queue_;
    // a request at most this large may be merged with others; a merged batch stays below the staging limits
    static constexpr uint64_t kSmallSrc = 4ull << 20, kSmallDst = 16ull << 20;
    static constexpr uint64_t kStageSrc = 32ull << 20, kStageDst = 128ull << 20;

    static bool small(const Req* r) { return r->src_bytes <= kSmallSrc && r->dst_bytes <= kSmallDst; }

    // decodes one request straight from / into its own buffers
    static void run_solo(kzg_ctx* ctx, Req* r) {
        int rc = kzg_ctx_set_dictionary(ctx, r->dict->empty() ? nullptr : r->dict->data(), r->dict->size());
        if (rc == 0)
            rc = kzg_decompress_batch_host(ctx, r->src, r->off, r->size, r->dst, r->doff, r->content, r->n, r->status, r->osize);
        r->rc = rc;
    }

    // gathers the requests' frames into pinned staging and decodes them as one batch
    static void run_merged(kzg_ctx* ctx, std::vector<Req*>& take, Gen* g) {
        uint64_t n = 0, sb = 0, db = 0;
        for (Req* r : take) { n += r->n; sb += r->src_bytes + 16; db += r->dst_bytes; }
        int rc = 0;
        if (g->src.cap < sb + 64) { g_pool.release(g->src); g->src = g_pool.acquire(sb + 64); }
        if (g->dst.cap < db + 64) { g_pool.release(g->dst); g->dst = g_pool.acquire(db + 64); }
        if (!g->src.p || !g->dst.p) rc = -(int)KZG_E_MEMORY;
        if (rc == 0) {
            try {
                g->off.resize(n); g->size.resize(n); g->doff.resize(n); g->cap.resize(n); g->osize.resize(n); g->status.resize(n);
            } catch (const std::bad_alloc&) { rc = -(int)KZG_E_MEMORY; }
        }
        if (rc == 0) {
            uint64_t k = 0, so = 0, dbase = 0;
            for (Req* r : take) {
                r->gen_base = dbase;
                for (size_t i = 0; i < r->n; i++, k++) {
                    memcpy(g->src.p + so, r->src + r->off[i], r->size[i]);
                    g->off[k] = so; g->size[k] = r->size[i];
                    g->doff[k] = dbase + r->doff[i]; g->cap[k] = r->content[i];
                    so += r->size[i];
                }
                so = (so + 15) & ~15ull;
                dbase += r->dst_bytes;
            }
            const std::vector<u8>* dict = take[0]->dict;
            rc = kzg_ctx_set_dictionary(ctx, dict->empty() ? nullptr : dict->data(), dict->size());
            if (rc == 0)
                rc = kzg_decompress_batch_host(ctx, g->src.p, g->off.data(), g->size.data(), g->dst.p, g->doff.data(), g->cap.data(), n,
                                               g->status.data(), g->osize.data());
        }
        uint64_t k = 0;
        for (Req* r : take) {
            r->rc = rc;
            if (rc == 0) {
                memcpy(r->status, g->status.data() + k, r->n * sizeof(uint32_t));
                memcpy(r->osize, g->osize.data() + k, r->n * sizeof(uint64_t));
                r->gen = g;
            }
            k += r->n;
        }
    }

public:
    bool ready() {
        std::lock_guard<std::mutex> lock(mu_);
        if (!lanes_[0].ctx) {
            kzg_ctx* c = nullptr;
            if (kzg_ctx_create(&c, -1) != 0) return false;
            lanes_[0].ctx = c;
        }
        return true;
    }

    int submit(Req& r) {
        std::unique_lock<std::mutex> lk(mu_);
        if (!lanes_[0].ctx) return -(int)KZG_E_MEMORY;
        queue_.push_back(&r);
        while (!r.done) {
            Lane* lane = nullptr;
            for (int i = 0; i < kLanes && !lane && !r.taken; i++)
                if (!lanes_[i].busy) lane = &lanes_[i];
            if (!lane) { cv_.wait(lk); continue; }          // in a running batch, or every lane is busy
            lane->busy = true;                              // this thread leads the next batch on this lane
            std::vector<Req*> take;
            Req* first = queue_.front();
            Gen* g = nullptr;
            if (!small(first)) { take.push_back(first); first->taken = true; queue_.erase(queue_.begin()); }
            else {
                uint64_t sb = 0, db = 0;
                size_t w = 0;
                for (size_t i = 0; i < queue_.size(); i++) {
                    Req* q = queue_[i];
                    if (small(q) && *q->dict == *first->dict && sb + q->src_bytes + 16 <= kStageSrc && db + q->dst_bytes <= kStageDst) {
                        take.push_back(q); sb += q->src_bytes + 16; db += q->dst_bytes;
                    } else queue_[w++] = q;
                }
                queue_.resize(w);
                for (Req* q : take) q->taken = true;
                if (take.size() > 1) {
                    g = &lane->gens[lane->flip++ & 1];
                    while (g->readers) cv_.wait(lk);        // readers of this lane's batch before last; they never wait on us
                }
            }
            lk.unlock();
            int rc = 0;
            if (!lane->ctx) {                               // further lanes come into being on first use
                kzg_ctx* c = nullptr;
                if (kzg_ctx_create(&c, -1) == 0) lane->ctx = c; else rc = -(int)KZG_E_MEMORY;
            }
            if (rc < 0) for (Req* q : take) q->rc = rc;
            else if (g) run_merged(lane->ctx, take, g);
            else run_solo(lane->ctx, take[0]);
            lk.lock();
            for (Req* q : take) { if (q->gen) q->gen->readers++; q->done = true; }
            lane->busy = false;
            cv_.notify_all();
        }
        lk.unlock();
        if (r.gen) {                                        // merged run: fetch this request's frames from staging
            for (size_t i = 0; i < r.n; i++)
                if (r.status[i] == 0) memcpy(r.dst + r.doff[i], r.gen->dst.p + r.gen_base + r.doff[i], r.osize[i]);
            lk.lock();
            if (--r.gen->readers == 0) cv_.notify_all();
            lk.unlock();
        }
        return r.rc;
    }
};

Batcher g_batcher;

}  // namespace

struct kzg_DCtx {
    std::vector<u8> in;          // compressed bytes of a frame that is not complete yet
    Arena out;                   // decoded bytes not yet handed to the caller: out.p[out_pos, out_size)
    size_t out_size = 0, out_pos = 0;
    std::vector<u8> dict;
    std::vector<uint64_t> off, size, content, doff;
    std::vector<uint32_t> status;
    std::vector<uint64_t> osize;
    size_t sticky_error = 0;
    size_t deferred_error = 0;   // error of a later frame of the same call: reported once the output of the frames before it is delivered
};

// Most output one call may hold (all frames completed by that call together); beyond it the call fails with
// memory_allocation instead of reserving whatever a few bytes of RLE blocks can claim.  KZG_STREAM_MAX_OUT_MB overrides.
static uint64_t stream_max_out() {
    static const uint64_t v = [] { const char* e = getenv("KZG_STREAM_MAX_OUT_MB"); const uint64_t m = e ? strtoull(e, nullptr, 10) : 0; return (m ? m : 4096ull) << 20; }();
    return v;
}

extern "C" kzg_DCtx* kzg_createDCtx(void) {
    if (!g_batcher.ready()) return nullptr;       // no GPU: no decoder (there is no CPU fallback)
    return new (std::nothrow) kzg_DCtx();
}

extern "C" size_t kzg_freeDCtx(kzg_DCtx* d) {
    if (d) g_pool.release(d->out);
    delete d;
    return 0;
}

extern "C" size_t kzg_DCtx_loadDictionary(kzg_DCtx* d, const void* dict, size_t n) {
    if (!d) return err(KZG_E_GENERIC);
    if (!d->in.empty() || d->out_pos < d->out_size) return err(60 /* stage_wrong */);
    d->dict.assign((const u8*)dict, (const u8*)dict + (dict ? n : 0));
    return 0;
}

// Decodes the complete frames in [p, p+len); appends output to d->out. Returns bytes consumed or error.
static size_t decode_complete_frames(kzg_DCtx* d, const u8* p, size_t len, size_t* n_frames_out) {
    *n_frames_out = 0;
    size_t consumed = 0;
    const size_t cap = kzg_scan_frames(p, len, nullptr, nullptr, nullptr, (size_t)-1, &consumed);     // count, then fill
    if (kzg_isError(cap)) return cap;
    if (cap == 0) return 0;
    try { d->off.resize(cap); d->size.resize(cap); d->content.resize(cap); } catch (const std::bad_alloc&) { return err(KZG_E_MEMORY); }
    const size_t n = kzg_scan_frames(p, len, d->off.data(), d->size.data(), d->content.data(), cap, &consumed);
    if (kzg_isError(n)) return n;
    if (n == 0) return 0;
    try { d->doff.resize(n); d->status.resize(n); d->osize.resize(n); } catch (const std::bad_alloc&) { return err(KZG_E_MEMORY); }
    // ZSTD_decompressStream refuses frames whose window exceeds ZSTD_WINDOWLOG_LIMIT_DEFAULT (2^27)
    // (windowLog 28 -> -16); through Kompressor's small output slices that
    // also holds for single-segment frames.
    for (size_t i = 0; i < n; i++) {
        KzgFrameHeader h;
        if (d->size[i] >= 5 && kzg_rd32(p + d->off[i]) == KZG_MAGIC &&
            kzg_parse_frame_header(p + d->off[i], d->size[i], h) == KZG_OK && h.window_size > (1ull << 27) + 1)
            return err(KZG_E_WINDOW_TOO_LARGE);
    }
    uint64_t total = 0, src_total = 0;
    for (size_t i = 0; i < n; i++) {
        // (content[i] is bounded by what the frame's blocks can regenerate, see kzg_scan_frames: no sum here can wrap)
        d->doff[i] = total; total += (d->content[i] + 15) & ~15ull; src_total += d->size[i];
        if (total > stream_max_out()) return err(KZG_E_MEMORY);
    }
    // output arena: what is still pending moves to the front of a buffer that also holds the new frames
    const size_t pending = d->out_size - d->out_pos;
    if (pending == 0) { d->out_pos = d->out_size = 0; }
    if (d->out.cap < pending + total + 64 || d->out_pos) {
        if (d->out.cap >= pending + total + 64) {
            memmove(d->out.p, d->out.p + d->out_pos, pending);
        } else {
            Arena a = g_pool.acquire(pending + total + 64);
            if (!a.p) return err(KZG_E_MEMORY);
            if (pending) memcpy(a.p, d->out.p + d->out_pos, pending);
            g_pool.release(d->out);
            d->out = a;
        }
        d->out_pos = 0;
        d->out_size = pending;
    }
    const size_t base = d->out_size;
    Req r;
    r.src = p; r.off = d->off.data(); r.size = d->size.data(); r.content = d->content.data(); r.doff = d->doff.data();
    r.n = n; r.src_bytes = src_total; r.dst_bytes = total;
    r.dst = d->out.p + base; r.status = d->status.data(); r.osize = d->osize.data(); r.dict = &d->dict;
    const int rc = g_batcher.submit(r);
    if (rc < 0) return err((u32)(-rc));
    // compact the outputs (frames were decoded at capacity-spaced offsets)
    size_t w = base;
    for (size_t i = 0; i < n; i++) {
        if (d->status[i]) {
            // libzstd has delivered the frames in front of a bad one by the time it reports the error: keep their
            // output and report the error when it has been handed over
            d->out_size = w;
            *n_frames_out = i;
            if (w == base && d->out_pos == d->out_size) return err(d->status[i]);
            d->deferred_error = err(d->status[i]);
            return consumed;
        }
        if (w != base + d->doff[i]) memmove(d->out.p + w, d->out.p + base + d->doff[i], d->osize[i]);
        w += d->osize[i];
    }
    d->out_size = w;
    *n_frames_out = n;
    return consumed;
}

extern "C" size_t kzg_decompressStream(kzg_DCtx* d, kzg_outBuffer* output, kzg_inBuffer* input) {
    if (!d || !output || !input) return err(KZG_E_GENERIC);
    if (input->pos > input->size || output->pos > output->size) return err(KZG_E_SRCSIZE_WRONG);
    if (d->sticky_error) return d->sticky_error;
    const u8* src = (const u8*)input->src + input->pos;
    const size_t avail = input->size - input->pos;
    size_t frames_done = 0;
    if (avail || !d->in.empty()) {
        size_t used;
        if (d->in.empty()) {                      // fast path: decode straight from the caller's buffer
            used = decode_complete_frames(d, src, avail, &frames_done);
            if (kzg_isError(used)) { d->sticky_error = used; return used; }
            d->in.assign(src + used, src + avail);
        } else {
            d->in.insert(d->in.end(), src, src + avail);
            used = decode_complete_frames(d, d->in.data(), d->in.size(), &frames_done);
            if (kzg_isError(used)) { d->sticky_error = used; return used; }
            d->in.erase(d->in.begin(), d->in.begin() + used);
        }
        input->pos = input->size;                 // zstd never leaves input behind (input.insufficient is never set)
    }
    size_t delivered = 0;
    {
        const size_t have = d->out_size - d->out_pos, room = output->size - output->pos;
        delivered = have < room ? have : room;
        if (delivered) memcpy((u8*)output->dst + output->pos, d->out.p + d->out_pos, delivered);
        output->pos += delivered;
        d->out_pos += delivered;
        if (d->out_pos == d->out_size) {
            d->out_pos = d->out_size = 0;
            if (d->out.cap > (1u << 20)) g_pool.release(d->out);      // large arenas go back to the pool at once
        }
    }
    const size_t pending = d->out_size - d->out_pos;
    if (pending) return pending;                          // undelivered output: caller must come back with room
    if (d->deferred_error) { d->sticky_error = d->deferred_error; d->in.clear(); return d->sticky_error; }
    if (!d->in.empty()) return d->in.size() < 9 ? 9 - d->in.size() : 3;   // inside a frame: hint for more input
    if (frames_done || delivered) return 0;               // frame boundary, everything flushed
    return 9;                                             // idle at a frame start (libzstd: header + block header hint)
}

// ------------------------------------------------------------------------------------------------
// Compress side of the shim: the libzstd functions behind
// Java_..._ZstdWrapper_createCompressor / freeCompressor / setParameter / loadCompressorDictionary / compressStream
// (Wrapper.cpp:10-56,75-121; Kotlin: ZstdCompressor.jvm.kt:10-42), on top of the batch frame writer.
// The stream is cut into chunks of 64 KiB and every chunk becomes one independent frame, so what this
// compressor writes is what the batch decoder reads fastest; any zstd decoder accepts the concatenation.
// In transform(ByteArray) the first call offers the whole input with finish = true (SliceTransform.kt:33-45):
// all chunks are compressed as one GPU batch and the output is drained over the following calls.
namespace {
std::mutex g_enc_mu;
kzg_enc* g_enc = nullptr;
const size_t kChunk = 65536;
}  // namespace

struct kzg_CCtx {
    std::vector<u8> in;              // input not yet compressed (less than one chunk between calls)
    Arena out;                       // frames not yet handed to the caller: out.p[out_pos, out_size)
    size_t out_size = 0, out_pos = 0;
    bool checksum = false;
    bool lazy = true;                // compression level >= 3 (the reference's default)
    bool wrote_frame = false;        // since the last completed e_end
    std::vector<uint64_t> off, size, doff, cap, osize;
    std::vector<uint32_t> status;
};

extern "C" kzg_CCtx* kzg_createCCtx(void) {
    {
        std::lock_guard<std::mutex> lock(g_enc_mu);
        if (!g_enc && kzg_enc_create(&g_enc, -1) != 0) return nullptr;      // no GPU: no compressor
    }
    return new (std::nothrow) kzg_CCtx();
}

extern "C" size_t kzg_freeCCtx(kzg_CCtx* c) {
    if (c) g_pool.release(c->out);
    delete c;
    return 0;
}

extern "C" size_t kzg_CCtx_setParameter(kzg_CCtx* c, int param, int value) {
    if (!c) return err(KZG_E_GENERIC);
    if (param == 201) { c->checksum = value != 0; return 0; }               // ZSTD_c_checksumFlag
    if (param == 100) { c->lazy = value >= 3; return 0; }                   // ZSTD_c_compressionLevel: greedy below 3, lazy from 3
    // window / hash / chain / search logs, minMatch, targetLength, strategy, content-size and dictionary-ID flags:
    // accepted without effect; the frame writer always writes the content size
    if ((param >= 101 && param <= 107) || param == 200 || param == 202) return 0;
    return err(40 /* parameter_unsupported */);
}

extern "C" size_t kzg_CCtx_loadDictionary(kzg_CCtx* c, const void* dict, size_t n) {
    if (!c) return err(KZG_E_GENERIC);
    if (!dict || n == 0) return 0;
    // A raw-content dictionary only adds history: frames that do not reference it stay valid for a decoder that has
    // it loaded.  A formatted one would also change the decoder's initial repeat offsets, which this writer relies on.
    if (n >= 8 && kzg_rd32((const u8*)dict) == KZG_DICT_MAGIC) return err(40 /* parameter_unsupported */);
    return 0;
}

// Compresses the first `nbytes` of c->in (whole chunks, plus a short last one) and appends the frames to c->out.
static size_t compress_buffered(kzg_CCtx* c, size_t nbytes) {
    const size_t n = nbytes ? (nbytes + kChunk - 1) / kChunk : 1;            // nbytes == 0: the empty frame
    try {
        c->off.resize(n); c->size.resize(n); c->doff.resize(n); c->cap.resize(n); c->osize.resize(n); c->status.resize(n);
    } catch (const std::bad_alloc&) { return err(KZG_E_MEMORY); }
    uint64_t total_cap = 0;
    for (size_t i = 0; i < n; i++) {
        c->off[i] = i * kChunk;
        c->size[i] = nbytes - i * kChunk < kChunk ? nbytes - i * kChunk : kChunk;
        c->cap[i] = (kzg_compress_bound((size_t)c->size[i]) + 15) & ~(size_t)15;
        c->doff[i] = total_cap;
        total_cap += c->cap[i];
    }
    const size_t pending = c->out_size - c->out_pos;
    Arena a = g_pool.acquire(pending + total_cap + 64);
    if (!a.p) return err(KZG_E_MEMORY);
    if (pending) memcpy(a.p, c->out.p + c->out_pos, pending);
    g_pool.release(c->out);
    c->out = a;
    c->out_pos = 0;
    c->out_size = pending;
    int rc;
    {
        std::lock_guard<std::mutex> lock(g_enc_mu);
        rc = kzg_compress_batch_host(g_enc, c->in.data(), c->off.data(), c->size.data(), c->out.p + pending, c->doff.data(),
                                     c->cap.data(), n, (c->checksum ? KZG_COMPRESS_CHECKSUM : 0) | (c->lazy ? KZG_COMPRESS_LAZY : 0), c->status.data(), c->osize.data());
    }
    if (rc < 0) return err((u32)(-rc));
    size_t w = pending;
    for (size_t i = 0; i < n; i++) {
        if (c->status[i]) return err(c->status[i]);
        if (w != pending + c->doff[i]) memmove(c->out.p + w, c->out.p + pending + c->doff[i], c->osize[i]);
        w += c->osize[i];
    }
    c->out_size = w;
    c->wrote_frame = true;
    return 0;
}

extern "C" size_t kzg_compressStream2(kzg_CCtx* c, kzg_outBuffer* output, kzg_inBuffer* input, int end_op) {
    if (!c || !output || !input) return err(KZG_E_GENERIC);
    if (input->pos > input->size || output->pos > output->size) return err(KZG_E_SRCSIZE_WRONG);
    if (end_op < 0 || end_op > 2) return err(42 /* parameter_outOfBound */);
    const size_t avail = input->size - input->pos;
    if (avail) {
        try { c->in.insert(c->in.end(), (const u8*)input->src + input->pos, (const u8*)input->src + input->size); }
        catch (const std::bad_alloc&) { return err(KZG_E_MEMORY); }
        input->pos = input->size;
    }
    // whole chunks are compressed as soon as they are buffered; flush / end also take the short rest
    size_t take = end_op == 0 ? c->in.size() / kChunk * kChunk : c->in.size();
    if (take || (end_op == 2 && !c->wrote_frame && c->out_size == c->out_pos)) {
        const size_t e = compress_buffered(c, take);
        if (kzg_isError(e)) return e;
        c->in.erase(c->in.begin(), c->in.begin() + take);
    }
    const size_t have = c->out_size - c->out_pos, room = output->size - output->pos;
    const size_t n = have < room ? have : room;
    if (n) memcpy((u8*)output->dst + output->pos, c->out.p + c->out_pos, n);
    output->pos += n;
    c->out_pos += n;
    const size_t pending = c->out_size - c->out_pos;
    if (pending == 0) {
        c->out_pos = c->out_size = 0;
        if (c->out.cap > (1u << 20)) g_pool.release(c->out);
        if (end_op == 2) c->wrote_frame = false;               // the stream is complete: the next call starts a new one
    }
    return pending;
}
