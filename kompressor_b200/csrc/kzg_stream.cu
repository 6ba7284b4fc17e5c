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
#include "../../include/kzg.h"
#include "kzg_parse.cuh"

#include <mutex>
#include <new>
#include <string.h>
#include <vector>



namespace {
std::mutex g_mu;                 // one shared device context per process; batches are serialised
kzg_ctx* g_ctx = nullptr;

kzg_ctx* shared_ctx() {
    if (!g_ctx) {
        kzg_ctx* c = nullptr;
        if (kzg_ctx_create(&c, -1) != 0) return nullptr;
        g_ctx = c;
    }
    return g_ctx;
}
inline size_t err(u32 code) { return (size_t)0 - (size_t)code; }
}  // namespace

struct kzg_DCtx {
    std::vector<u8> in;          // compressed bytes of a frame that is not complete yet
    std::vector<u8> out;         // decoded bytes not yet handed to the caller
    size_t out_pos = 0;
    std::vector<u8> dict;
    std::vector<uint64_t> off, size, content, doff;
    std::vector<uint32_t> status;
    std::vector<uint64_t> osize;
    size_t sticky_error = 0;
};

extern "C" kzg_DCtx* kzg_createDCtx(void) {
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (!shared_ctx()) return nullptr;        // no GPU: no decoder (there is no CPU fallback)
    }
    return new (std::nothrow) kzg_DCtx();
}

extern "C" size_t kzg_freeDCtx(kzg_DCtx* d) {
    delete d;
    return 0;
}

extern "C" size_t kzg_DCtx_loadDictionary(kzg_DCtx* d, const void* dict, size_t n) {
    if (!d) return err(KZG_E_GENERIC);
    if (!d->in.empty() || d->out_pos < d->out.size()) return err(60 /* stage_wrong */);
    d->dict.assign((const u8*)dict, (const u8*)dict + (dict ? n : 0));
    return 0;
}

// Decodes the complete frames in [p, p+len); appends output to d->out. Returns bytes consumed or error.
static size_t decode_complete_frames(kzg_DCtx* d, const u8* p, size_t len, size_t* n_frames_out) {
    size_t cap = len / 9 + 2;
    d->off.resize(cap); d->size.resize(cap); d->content.resize(cap);
    size_t consumed = 0;
    const size_t n = kzg_scan_frames(p, len, d->off.data(), d->size.data(), d->content.data(), cap, &consumed);
    *n_frames_out = 0;
    if (kzg_isError(n)) return n;
    if (n == 0) return 0;
    d->doff.resize(n); d->status.resize(n); d->osize.resize(n);
    // ZSTD_decompressStream refuses frames whose window exceeds ZSTD_WINDOWLOG_LIMIT_DEFAULT (2^27)
    // (windowLog 28 -> -16); through Kompressor's small output slices that
    // also holds for single-segment frames.
    for (size_t i = 0; i < n; i++) {
        KzgFrameHeader h;
        if (d->size[i] >= 5 && kzg_rd32(p + d->off[i]) == KZG_MAGIC &&
            kzg_parse_frame_header(p + d->off[i], d->size[i], h) == KZG_OK && h.window_size > (1ull << 27) + 1)
            return err(KZG_E_WINDOW_TOO_LARGE);
    }
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) { d->doff[i] = total; total += (d->content[i] + 15) & ~15ull; }
    const size_t base = d->out.size();
    try { d->out.resize(base + total + 64); } catch (const std::bad_alloc&) { return err(KZG_E_MEMORY); }
    int rc;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        kzg_ctx* c = shared_ctx();
        if (!c) return err(KZG_E_MEMORY);
        rc = kzg_ctx_set_dictionary(c, d->dict.empty() ? nullptr : d->dict.data(), d->dict.size());
        if (rc == 0)
            rc = kzg_decompress_batch_host(c, p, d->off.data(), d->size.data(), d->out.data() + base, d->doff.data(),
                                           d->content.data(), n, d->status.data(), d->osize.data());
    }
    if (rc < 0) { d->out.resize(base); return err((u32)(-rc)); }
    // compact the outputs (frames were decoded at capacity-spaced offsets)
    size_t w = base;
    for (size_t i = 0; i < n; i++) {
        if (d->status[i]) { d->out.resize(base); return err(d->status[i]); }
        if (w != base + d->doff[i]) memmove(d->out.data() + w, d->out.data() + base + d->doff[i], d->osize[i]);
        w += d->osize[i];
    }
    d->out.resize(w);
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
        const size_t have = d->out.size() - d->out_pos, room = output->size - output->pos;
        delivered = have < room ? have : room;
        if (delivered) memcpy((u8*)output->dst + output->pos, d->out.data() + d->out_pos, delivered);
        output->pos += delivered;
        d->out_pos += delivered;
        if (d->out_pos == d->out.size()) { d->out.clear(); d->out_pos = 0; }
    }
    const size_t pending = d->out.size() - d->out_pos;
    if (pending) return pending;                          // undelivered output: caller must come back with room
    if (!d->in.empty()) return d->in.size() < 9 ? 9 - d->in.size() : 3;   // inside a frame: hint for more input
    if (frames_done || delivered) return 0;               // frame boundary, everything flushed
    return 9;                                             // idle at a frame start (libzstd: header + block header hint)
}
