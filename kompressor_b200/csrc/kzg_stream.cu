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
#include <stdio.h>
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

    // Runs f(ctx) with one device context to itself (segment decoding of one handle).
    template <class F> int with_lane(F&& f) {
        std::unique_lock<std::mutex> lk(mu_);
        Lane* lane = nullptr;
        for (;;) {
            for (int i = 0; i < kLanes && !lane; i++) if (!lanes_[i].busy) lane = &lanes_[i];
            if (lane) break;
            cv_.wait(lk);
        }
        lane->busy = true;
        lk.unlock();
        int rc = 0;
        if (!lane->ctx) {
            kzg_ctx* c = nullptr;
            if (kzg_ctx_create(&c, -1) == 0) lane->ctx = c; else rc = -(int)KZG_E_MEMORY;
        }
        if (rc == 0) rc = f(lane->ctx);
        lk.lock();
        lane->busy = false;
        cv_.notify_all();
        return rc;
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

// ------------------------------------------------------------------------------------------------ decompress side
// The handle behaves like ZSTD_decompressStream, call for call (Wrapper.cpp:178; driven by SliceTransform.kt:33-45 and
// SliceTransformRawSource.kt:32-55).  What a caller of libzstd can observe, and therefore what is restated here:
//   * output leaves block by block; a call returns when the output slice is full, when the input runs dry (all of it has
//     been taken over then), or -- with 0 -- when a frame has been handed over completely, and then the input is only
//     consumed up to the end of that frame;
//   * an error is returned by the call that reaches it, and that call hands over nothing (libzstd does not advance the
//     positions on an error): what the caller keeps is what earlier calls delivered -- the frames in front of a bad one,
//     and of the bad frame the blocks in front of the bad block as far as whole calls delivered them;
//   * an incomplete frame is not an error: its complete blocks are decoded and handed over, the rest waits.
// Underneath, decoding runs AHEAD of what has been handed over, in rounds, so that the GPU sees batches: a round takes
// every complete frame at the parse position (up to kAhead bytes of output; one GPU batch through the cross-handle
// batcher above), or -- for a frame that is not complete yet or larger than a segment -- the next run of complete
// blocks of that frame (kzg_decompress_segment_host: entropy tables, repeat offsets, checksum and the last window
// bytes of output carry over).  A round starts only when everything decoded so far has been handed over, so a handle
// holds at most kAhead bytes of output plus the frame's window, whatever the size of the stream.
struct Unit {                    // one step of the hand-over: a frame, a segment of one, or an error
    size_t out_len = 0;          // bytes in the arena
    uint64_t in_end = 0;         // stream offset consumed through this unit
    bool frame_end = false;      // the call that finishes it returns 0
    size_t error = 0;            // nonzero: no data; the call that reaches it returns this
};

struct SegFrame {                // the frame that is being decoded in segments
    bool active = false, last_seen = false;      // last_seen: every block is decoded, only the checksum is missing
    kzg_seg_state st;
    uint64_t fcs = ~0ull, produced = 0, window = 0;
    bool has_checksum = false;
    u32 raw_left = 0;            // inside a raw block that arrives in pieces: bytes of it still to come
    bool raw_last = false;       // ... and whether it is the frame's last block
    std::vector<u8> hist;        // the last `window` bytes of content (in front of them: what is left of the dictionary)
};

struct kzg_DCtx {
    std::vector<u8> dict;
    uint64_t reported = 0;       // stream bytes the caller has been told are consumed
    uint64_t goal = 0;           // ... and those libzstd would have taken by now (told as far as the caller's slices reach)
    uint64_t engine = 0;         // stream bytes decoded (or skipped) so far; >= reported
    std::vector<u8> own;         // stream bytes [engine, engine + own.size()) taken over from the caller
    Arena out;                   // decoded bytes not yet handed over: out.p[out_pos, out_size)
    size_t out_size = 0, out_pos = 0;
    std::vector<Unit> units;     // hand-over queue, units[unit_head] is current
    size_t unit_head = 0, cur_left = 0;
    bool cur_open = false;       // cur_left is valid for units[unit_head]
    SegFrame seg;
    std::vector<uint64_t> off, size, content, doff;
    std::vector<uint32_t> status;
    std::vector<uint64_t> osize;
    size_t sticky_error = 0;
};

namespace {

const uint64_t kAhead = 64ull << 20;         // most output one round decodes ahead of the hand-over
const uint64_t kSegMax = 32ull << 20;        // most output of one segment; frames above it are decoded in segments

// Most output one handle may hold.  KZG_STREAM_MAX_OUT_MB overrides (tests).
uint64_t stream_max_out() {
    static const uint64_t v = [] { const char* e = getenv("KZG_STREAM_MAX_OUT_MB"); const uint64_t m = e ? strtoull(e, nullptr, 10) : 0; return (m ? m : 4096ull) << 20; }();
    return v;
}

bool arena_reserve(kzg_DCtx* d, size_t extra) {          // room for `extra` more bytes behind out_size
    if (d->out.cap >= d->out_size + extra + 64) return true;
    Arena a = g_pool.acquire(d->out_size + extra + 64);
    if (!a.p) return false;
    if (d->out_size) memcpy(a.p, d->out.p, d->out_size);
    g_pool.release(d->out);
    d->out = a;
    return true;
}

void push_unit(kzg_DCtx* d, size_t out_len, uint64_t in_end, bool frame_end, size_t error) {
    Unit u; u.out_len = out_len; u.in_end = in_end; u.frame_end = frame_end; u.error = error;
    d->units.push_back(u);
}

// Walks whole blocks in [v, v + n): how many, where the last header is, their output bound.  Stops at the last block of
// the frame, at an incomplete block, at `max_blocks`, or when the bound would exceed `max_out` (never before one block).
struct BlockRun { size_t bytes = 0, last_hdr = 0, count = 0; uint64_t bound = 0; bool last = false, starved = false; u32 error = 0; };
BlockRun walk_blocks(const u8* v, size_t n, u32 block_max, uint64_t max_out, size_t max_blocks) {
    BlockRun r;
    size_t q = 0;
    while (q + 3 <= n && r.count < max_blocks) {
        const u32 bh = kzg_rd24(v + q);
        const u32 type = (bh >> 1) & 3, bsz = bh >> 3;
        if (type == 3 || bsz > block_max) { r.error = KZG_E_CORRUPTION; break; }
        const size_t adv = type == 1 ? 1 : bsz;
        if (q + 3 + adv > n) { r.starved = true; break; }
        const uint64_t add = type == 2 ? block_max : bsz;
        if (r.count && r.bound + add > max_out) break;
        r.bound += add; r.last_hdr = q; r.count++;
        q += 3 + adv;
        r.bytes = q;
        if (bh & 1) { r.last = true; break; }
    }
    if (!r.last && !r.error && q + 3 > n && r.count < max_blocks) r.starved = true;
    return r;
}

// A raw block whose content is there only in part (at least one byte of it): libzstd passes raw bytes through as they
// arrive instead of waiting for the whole block.
bool partial_raw_at(const u8* p, size_t n, u32 block_max) {
    if (n < 4) return false;
    const u32 bh = kzg_rd24(p), type = (bh >> 1) & 3, bsz = bh >> 3;
    return type == 0 && bsz <= block_max && (size_t)bsz > n - 3;
}

// One segment of the active frame, decoded behind out_size.  Returns 0 or an error value; *osize / *status as the batch calls.
size_t run_segment(kzg_DCtx* d, const u8* blocks, const BlockRun& r, uint32_t* status, uint64_t* osize) {
    if (!arena_reserve(d, (size_t)r.bound)) return err(KZG_E_MEMORY);
    u8* dst = d->out.p + d->out_size;
    SegFrame& sg = d->seg;
    const int rc = g_batcher.with_lane([&](kzg_ctx* ctx) {
        return kzg_decompress_segment_host(ctx, &sg.st, blocks, r.bytes, r.last_hdr, sg.hist.data(), sg.hist.size(), dst, (size_t)r.bound + 32, status, osize);
    });
    if (rc < 0) return err((u32)(-rc));
    return 0;
}

void seg_commit(kzg_DCtx* d, uint64_t n) {                // n bytes behind out_size are content of the active frame
    SegFrame& sg = d->seg;
    const u8* p = d->out.p + d->out_size;
    if (n >= sg.window) sg.hist.assign(p + (n - sg.window), p + n);
    else {
        sg.hist.insert(sg.hist.end(), p, p + n);
        if (sg.hist.size() > sg.window) sg.hist.erase(sg.hist.begin(), sg.hist.begin() + (sg.hist.size() - (size_t)sg.window));
    }
    sg.produced += n;
    d->out_size += (size_t)n;
}

// The blocks [0, k) of a run decoded on their own (what libzstd has handed over before it meets the bad block k), from
// the state `st0` the run started with.  Queues them as one unit.
void queue_good_prefix(kzg_DCtx* d, const u8* blocks, size_t n, u32 block_max, size_t k, const kzg_seg_state& st0, uint64_t in_end_base) {
    if (k == 0) return;
    d->seg.st = st0;
    const BlockRun r = walk_blocks(blocks, n, block_max, ~0ull, k);
    if (r.count != k || r.error) return;
    uint32_t st = 0; uint64_t os = 0;
    if (run_segment(d, blocks, r, &st, &os) != 0 || st != 0) return;
    seg_commit(d, os);
    push_unit(d, (size_t)os, in_end_base + r.bytes, false, 0);
}

// Segment mode: the next run of whole blocks of the active frame (or its checksum).  Returns the stream bytes taken.
size_t seg_round(kzg_DCtx* d, const u8* v, size_t n, bool* incomplete) {
    SegFrame& sg = d->seg;
    if (sg.last_seen) {
        if (n < 4) { *incomplete = true; return 0; }
        const bool ok = kzg_rd32(v) == kzg_segment_checksum(&sg.st);
        sg.active = false;
        if (ok) push_unit(d, 0, d->engine + 4, true, 0);
        else push_unit(d, 0, d->engine, false, err(KZG_E_CHECKSUM_WRONG));
        return ok ? 4 : 0;
    }
    if (sg.raw_left) {
        // the next piece of a raw block, wrapped as a raw block of its own so that it takes the same way as everything else
        // (history, running checksum)
        if (n == 0) { *incomplete = true; return 0; }
        const u32 take = n < sg.raw_left ? (u32)n : sg.raw_left;
        const bool ends_frame = take == sg.raw_left && sg.raw_last;
        std::vector<u8> piece(3 + (size_t)take);
        const u32 bh = take << 3;
        piece[0] = (u8)bh; piece[1] = (u8)(bh >> 8); piece[2] = (u8)(bh >> 16);
        memcpy(piece.data() + 3, v, take);
        BlockRun r1; r1.bytes = piece.size(); r1.last_hdr = 0; r1.count = 1; r1.bound = take; r1.last = ends_frame;
        uint32_t st = 0; uint64_t os = 0;
        const size_t e = run_segment(d, piece.data(), r1, &st, &os);
        if (!e && st == 0 && ends_frame && sg.fcs != ~0ull && sg.produced + os != sg.fcs) st = KZG_E_CORRUPTION;
        if (e || st) { sg.active = false; push_unit(d, 0, d->engine, false, e ? e : err(st)); return 0; }
        seg_commit(d, os);
        sg.raw_left -= take;
        push_unit(d, (size_t)os, d->engine + take, ends_frame && !sg.has_checksum, 0);
        if (ends_frame) { if (sg.has_checksum) sg.last_seen = true; else sg.active = false; }
        return take;
    }
    BlockRun r = walk_blocks(v, n, sg.st.block_max, kSegMax, ~(size_t)0);
    // libzstd compares the content size when it decodes the last block, before handing it over: that block goes alone
    if (r.last && r.count > 1 && sg.fcs != ~0ull) r = walk_blocks(v, n, sg.st.block_max, kSegMax, r.count - 1);
    // The frame is still arriving: libzstd takes its input block by block, and a pump that offers more input only when
    // everything has been taken sees that, so the blocks are handed over (and their input reported taken) one at a time.
    // A frame whose end is in sight goes in runs of up to 32 MiB.
    else if (r.starved && r.count > 1) r = walk_blocks(v, n, sg.st.block_max, kSegMax, 1);
    if (r.count == 0) {
        if (r.error) { sg.active = false; push_unit(d, 0, d->engine, false, err(r.error)); return 0; }
        if (partial_raw_at(v, n, sg.st.block_max)) {       // its header is taken now, its bytes piece by piece
            const u32 bh = kzg_rd24(v);
            sg.raw_left = bh >> 3; sg.raw_last = (bh & 1) != 0;
            return 3;
        }
        *incomplete = true;
        return 0;
    }
    const kzg_seg_state st0 = sg.st;
    const std::vector<u8> hist0 = sg.hist;         // (a failed run is decoded again up to the bad block)
    uint32_t st = 0; uint64_t os = 0;
    const size_t e = run_segment(d, v, r, &st, &os);
    if (getenv("KZG_STREAM_DEBUG")) fprintf(stderr, "seg_round: engine %llu n %zu blocks %zu bytes %zu bound %llu last %d -> rc %zd status %u osize %llu hist %zu\n",
                                            (unsigned long long)d->engine, n, r.count, r.bytes, (unsigned long long)r.bound, (int)r.last, (ssize_t)e, st, (unsigned long long)os, sg.hist.size());
    if (e) { sg.active = false; push_unit(d, 0, d->engine, false, e); return 0; }
    if (st == 0 && r.last && sg.fcs != ~0ull && sg.produced + os != sg.fcs) { st = KZG_E_CORRUPTION; os = 0; }
    if (st != 0) {
        sg.hist = hist0;
        queue_good_prefix(d, v, n, st0.block_max, (size_t)(os & 0x3FFFFFFFull), st0, d->engine);
        sg.active = false;
        // literals libzstd would have made up: with a checksum its verdict, at the end of the frame, is
        // checksum_wrong; the blocks in front of the first such block are handed over, the made-up ones are not
        if (((os >> 62) & 1) && sg.has_checksum) st = KZG_E_CHECKSUM_WRONG;
        push_unit(d, 0, d->engine, false, err(st));
        return 0;
    }
    seg_commit(d, os);
    const bool ends = r.last && !sg.has_checksum;
    push_unit(d, (size_t)os, d->engine + r.bytes, ends, 0);
    if (r.last) { if (sg.has_checksum) sg.last_seen = true; else sg.active = false; }
    return r.bytes;
}

// Starts segment mode for the frame at v (its header is complete).  Returns the header bytes taken, 0 with a queued error.
size_t seg_start(kzg_DCtx* d, const KzgFrameHeader& h) {
    SegFrame& sg = d->seg;
    u32 dict_id = 0;
    if (d->dict.size() >= 8 && kzg_rd32(d->dict.data()) == KZG_DICT_MAGIC) dict_id = kzg_rd32(d->dict.data() + 4);
    if (h.dict_id && h.dict_id != dict_id) { push_unit(d, 0, d->engine, false, err(KZG_E_DICT_WRONG)); return 0; }
    const u32 block_max = h.window_size < KZG_BLOCK_MAX ? (u32)h.window_size : KZG_BLOCK_MAX;
    const long long content = kzg_segment_begin(&sg.st, block_max, (int)h.has_checksum, d->dict.empty() ? nullptr : d->dict.data(), d->dict.size());
    if (content < 0) { push_unit(d, 0, d->engine, false, err((u32)(-content))); return 0; }
    sg.active = true; sg.last_seen = false; sg.raw_left = 0; sg.raw_last = false;
    sg.fcs = h.fcs; sg.produced = 0; sg.window = h.window_size; sg.has_checksum = h.has_checksum != 0;
    sg.hist.assign(d->dict.end() - (size_t)content, d->dict.end());
    return h.hdr_size;
}

// One round of decoding at the parse position: v[0, n) are the stream bytes from d->engine on.  Queues units; returns the
// stream bytes taken.  *incomplete: nothing could be done because the next block / frame is not complete.
size_t engine_round(kzg_DCtx* d, const u8* v, size_t n, bool* incomplete) {
    *incomplete = false;
    if (d->seg.active) return seg_round(d, v, n, incomplete);
    // ---- at a frame start: every complete frame from here on, as one batch
    size_t pos = 0, nf = 0;
    uint64_t total = 0;
    d->off.clear(); d->size.clear(); d->content.clear();
    size_t stop_error = 0;
    bool start_segments = false;
    KzgFrameHeader seg_h;
    while (pos < n) {
        const u8* f = v + pos;
        const size_t rem = n - pos;
        if (rem < 4) break;
        const u32 magic = kzg_rd32(f);
        if ((magic & 0xFFFFFFF0u) == KZG_SKIP_MAGIC) {
            if (rem < 8 || 8ull + kzg_rd32(f + 4) > rem) break;
            d->off.push_back(pos); d->size.push_back(8ull + kzg_rd32(f + 4)); d->content.push_back(0);
            pos += (size_t)d->size.back(); nf++;
            continue;
        }
        KzgFrameHeader h;
        const u32 e = magic != KZG_MAGIC ? (u32)KZG_E_PREFIX_UNKNOWN : kzg_parse_frame_header(f, rem, h);
        if (e == KZG_E_SRCSIZE_WRONG) break;
        if (e) { stop_error = err(e); break; }
        // ZSTD_decompressStream refuses frames whose window exceeds ZSTD_WINDOWLOG_LIMIT_DEFAULT (2^27) (SURVEY.md
        // Appendix C: windowLog 28 -> -16); through Kompressor's small output slices that also holds for single-segment frames
        if (h.window_size > (1ull << 27) + 1) { stop_error = err(KZG_E_WINDOW_TOO_LARGE); break; }
        const u32 block_max = h.window_size < KZG_BLOCK_MAX ? (u32)h.window_size : KZG_BLOCK_MAX;
        const BlockRun r = walk_blocks(f + h.hdr_size, rem - h.hdr_size, block_max, ~0ull, ~(size_t)0);
        const size_t fsz = h.hdr_size + r.bytes + (h.has_checksum ? 4 : 0);
        const bool complete = r.last && !r.error && fsz <= rem;
        if (!complete || r.bound > kSegMax) {
            // Not complete yet, broken block structure, or too large for one piece: block by block -- but only when a
            // whole block is there to hand over.  A frame whose first block is still arriving just waits: once it is
            // complete it goes through the batch path, where the calls of many handles are merged into one GPU batch.
            if (nf == 0 && (r.count || r.error || partial_raw_at(f + h.hdr_size, rem - h.hdr_size, block_max))) { start_segments = true; seg_h = h; }
            break;
        }
        // (the Frame_Content_Size field is untrusted: a frame cannot regenerate more than its blocks allow)
        const uint64_t cap = (h.fcs != ~0ull && h.fcs < r.bound) ? h.fcs : r.bound;
        if (nf && total + cap > kAhead) break;
        d->off.push_back(pos); d->size.push_back(fsz); d->content.push_back(cap);
        total += (cap + 15) & ~15ull;
        pos += fsz; nf++;
    }
    if (nf == 0) {
        if (start_segments) return seg_start(d, seg_h);
        if (stop_error) { push_unit(d, 0, d->engine, false, stop_error); return 0; }
        *incomplete = true;
        return 0;
    }
    if (total > stream_max_out()) { push_unit(d, 0, d->engine, false, err(KZG_E_MEMORY)); return 0; }
    try { d->doff.resize(nf); d->status.resize(nf); d->osize.resize(nf); } catch (const std::bad_alloc&) { push_unit(d, 0, d->engine, false, err(KZG_E_MEMORY)); return 0; }
    uint64_t t = 0, src_total = 0;
    for (size_t i = 0; i < nf; i++) { d->doff[i] = t; t += (d->content[i] + 15) & ~15ull; src_total += d->size[i]; }
    if (!arena_reserve(d, (size_t)total)) { push_unit(d, 0, d->engine, false, err(KZG_E_MEMORY)); return 0; }
    const size_t base = d->out_size;
    Req rq;
    rq.src = v; rq.off = d->off.data(); rq.size = d->size.data(); rq.content = d->content.data(); rq.doff = d->doff.data();
    rq.n = nf; rq.src_bytes = src_total; rq.dst_bytes = total;
    rq.dst = d->out.p + base; rq.status = d->status.data(); rq.osize = d->osize.data(); rq.dict = &d->dict;
    const int rc = g_batcher.submit(rq);
    if (rc < 0) { push_unit(d, 0, d->engine, false, err((u32)(-rc))); return 0; }
    // compact the outputs (frames were decoded at capacity-spaced offsets) and queue them
    size_t w = base, taken = 0;
    for (size_t i = 0; i < nf; i++) {
        const uint64_t in_end = d->engine + d->off[i] + d->size[i];
        if (d->status[i] == 0 || (d->status[i] == KZG_E_CHECKSUM_WRONG && !(d->osize[i] >> 63) && d->osize[i] <= d->content[i])) {
            if (w != base + d->doff[i]) memmove(d->out.p + w, d->out.p + base + d->doff[i], (size_t)d->osize[i]);
            w += (size_t)d->osize[i];
            d->out_size = w;
            if (d->status[i]) {      // the content is all there, the checksum does not match it: handed over, then the error
                push_unit(d, (size_t)d->osize[i], d->engine + d->off[i], false, 0);
                push_unit(d, 0, d->engine + d->off[i], false, err(d->status[i]));
                return taken;
            }
            push_unit(d, (size_t)d->osize[i], in_end, true, 0);
            taken = (size_t)(d->off[i] + d->size[i]);
            continue;
        }
        // a bad frame: what libzstd hands over first are the blocks in front of the bad one (out_size names it)
        d->out_size = w;
        const u8* f = v + d->off[i];
        KzgFrameHeader h;
        const size_t bad_block = (size_t)(d->osize[i] & 0x3FFFFFFFull);
        if ((kzg_rd32(f) & 0xFFFFFFF0u) != KZG_SKIP_MAGIC && kzg_parse_frame_header(f, (size_t)d->size[i], h) == KZG_OK && bad_block > 0 &&
            d->status[i] != KZG_E_DICT_WRONG) {
            d->engine += d->off[i];                          // (seg_start / the prefix units count from the frame's start)
            const size_t hs = seg_start(d, h);
            if (hs) {
                const kzg_seg_state st0 = d->seg.st;
                d->engine += hs;
                queue_good_prefix(d, f + hs, (size_t)d->size[i] - hs, st0.block_max, bad_block, st0, d->engine);
                d->engine -= hs;
            }
            d->seg.active = false;
            d->engine -= d->off[i];
        }
        push_unit(d, 0, d->engine + d->off[i], false, err(d->status[i]));
        return taken;
    }
    return taken;
}

}  // namespace

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
    if (d->engine != d->reported || !d->own.empty() || d->unit_head < d->units.size() || d->seg.active) return err(60 /* stage_wrong */);
    d->dict.assign((const u8*)dict, (const u8*)dict + (dict ? n : 0));
    return 0;
}

extern "C" size_t kzg_decompressStream(kzg_DCtx* d, kzg_outBuffer* output, kzg_inBuffer* input) {
    if (!d || !output || !input) return err(KZG_E_GENERIC);
    if (input->pos > input->size || output->pos > output->size) return err(KZG_E_SRCSIZE_WRONG);
    if (d->sticky_error) return d->sticky_error;
    const u8* src = (const u8*)input->src + input->pos;           // stream offset d->reported
    const size_t avail = input->size - input->pos;
    const size_t out_entry = output->pos, in_entry = input->pos;
    const uint64_t rep_entry = d->reported;                       // stream offset of src[0]
    // bytes of this call the handle has not seen yet: those behind what it decoded ahead or took over earlier
    const uint64_t seen_end = d->engine + d->own.size();
    size_t fresh_lo = seen_end > rep_entry ? (size_t)(seen_end - rep_entry) : 0;
    if (fresh_lo > avail) fresh_lo = avail;                       // (a caller that presents less than before: nothing new)
    bool waiting = false;                                         // the next block / frame is not complete
    size_t result = 0;
    // input is reported taken up to d->goal, as far as this call's slice reaches (a pump may offer LESS than it offered
    // before: BufferSliceTransformHelper limits the input slice by the output bytes its reader still wants)
    auto report = [&]() {
        const uint64_t lim = rep_entry + avail, to = d->goal < lim ? d->goal : lim;
        if (to > d->reported) { input->pos += (size_t)(to - d->reported); d->reported = to; }
    };
    report();
    for (;;) {
        // ---- hand over
        if (d->unit_head < d->units.size()) {
            Unit& u = d->units[d->unit_head];
            if (u.error) {                                        // libzstd: positions do not advance on an error
                output->pos = out_entry;
                input->pos = in_entry;
                d->sticky_error = u.error;
                return u.error;
            }
            if (!d->cur_open) {
                // libzstd has taken a block's input when it starts to hand its output over
                d->cur_open = true; d->cur_left = u.out_len;
                if (u.in_end > d->goal) d->goal = u.in_end;
                report();
            }
            const size_t room = output->size - output->pos;
            const size_t k = d->cur_left < room ? d->cur_left : room;
            if (k) memcpy((u8*)output->dst + output->pos, d->out.p + d->out_pos, k);
            output->pos += k; d->out_pos += k; d->cur_left -= k;
            if (d->cur_left) { result = d->cur_left; break; }     // the slice is full in the middle of a block run
            const bool fe = u.frame_end;
            const uint64_t in_end = u.in_end;
            d->cur_open = false;
            d->unit_head++;
            if (d->unit_head == d->units.size()) {                // everything decoded so far is handed over
                d->units.clear(); d->unit_head = 0;
                d->out_pos = d->out_size = 0;
                if (d->out.cap > (1u << 20)) g_pool.release(d->out);      // large arenas go back to the pool at once
            }
            if (in_end > d->goal) d->goal = in_end;
            report();
            if (fe) { result = 0; goto done; }                     // a frame boundary: libzstd returns 0 and stops consuming
            continue;
        }
        // ---- decode ahead: one round at the parse position
        const u8* v; size_t n;
        if (d->own.empty()) {
            const size_t skip = (size_t)(d->engine - rep_entry);
            v = src + skip; n = avail > skip ? avail - skip : 0;
        } else {
            if (fresh_lo < avail) { d->own.insert(d->own.end(), src + fresh_lo, src + avail); fresh_lo = avail; }
            v = d->own.data(); n = d->own.size();
        }
        bool incomplete = false;
        const size_t took = n ? engine_round(d, v, n, &incomplete) : 0;
        if (!n) incomplete = true;
        d->engine += took;
        if (!d->own.empty()) d->own.erase(d->own.begin(), d->own.begin() + took);
        if (d->unit_head < d->units.size() || took) continue;
        (void)incomplete;
        // nothing more can be decoded: libzstd has taken all the input into its own buffer by now
        if (d->own.empty() && n) d->own.assign(v, v + n);
        waiting = true;
        break;
    }
    if (waiting) {
        const uint64_t all = rep_entry + avail;
        if (all > d->goal) d->goal = all;
        input->pos = input->size;
        d->reported = all;
        // hint: inside a frame, or idle at a frame start (libzstd: header + block header)
        result = (d->seg.active || !d->own.empty()) ? (d->own.size() < 9 ? 9 - d->own.size() : 3) : 9;
    }
done:
    return result;
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
