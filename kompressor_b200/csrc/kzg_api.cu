// C ABI, layer A (batch API): context, workspace, frame-table builder.  See include/kzg.h.
#include "../../include/kzg.h"
#include "kzg_kernels.cuh"
#include "kzg_parse.cuh"
#include "kzg_entropy.cuh"

#include <mutex>
#include <vector>
#include <new>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "kzg: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -(int)KZG_E_MEMORY; } } while (0)

#define KZG_MAX_CHUNKS 32                 // host-pointer API: sub-batches pipelined over the copy engines
#define KZG_EV_RING 32                   // profiled batches whose per-stage events are kept (kzg_stage_ms_history)
#define KZG_CHUNK_BYTES (128ull << 20)   // target decoded bytes per sub-batch: big enough that a sub-batch still fills the GPU

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    // Contents are not preserved.  On failure the old buffer (and its capacity) stays as it was, so that a later,
    // smaller request still works.
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        const size_t want = bytes + (bytes >> 3) + 256;
        if (want < bytes) return -1;
        void* q = nullptr;
        if (cudaMalloc(&q, want) != cudaSuccess) {
            cudaGetLastError();
            // one more try after giving the old buffer back (the request may only fit without it)
            if (p) { cudaFree(p); p = nullptr; cap = 0; }
            if (cudaMalloc(&q, want) != cudaSuccess) { cudaGetLastError(); return -1; }
        }
        if (p) cudaFree(p);
        p = q; cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Device constants: predefined FSE distributions, packed LL/ML code tables, and the entropy tables of the
// formatted dictionary currently loaded (if any).
struct KzgConsts { KzgFseDesc def; u32 codes[96]; KzgFseDesc dict_fse; KzgHufDesc dict_huf; };

struct kzg_ctx {
    int device = 0;
    int sm_count = 0;
    std::mutex mu;
    DevBuf frames, blocks, cblocks, fse, huf, hlist, hlist2, slist, blist, dlist, lit, seq, fsetab, ctr, consts;
    // sizing state (sticky across batches, grown on overflow)
    u64 block_cap = 0, cblock_cap = 0, lit_cap = 0, seq_cap = 0;
    u32 tile_bytes = 0;
    double lit_factor = 0.75, seq_div = 12.0, block_div = 32768.0;
    KzgCounters* h_ctr = nullptr;   // pinned, KZG_MAX_CHUNKS slots (slot 0 = the device-pointer API)
    cudaStream_t last_stream = nullptr;
    cudaStream_t own_stream = nullptr;                 // host-pointer API: decode stream
    cudaStream_t in_stream = nullptr, out_stream = nullptr;   // host-pointer API: H2D / D2H copy streams
    cudaEvent_t ev_in[KZG_MAX_CHUNKS] = {}, ev_cmp[KZG_MAX_CHUNKS] = {}, ev_tab = nullptr;
    u32 last_launches = 0;
    u64 last_n_frames = 0;
    // staging for the host-pointer form
    DevBuf st_src, st_dst, st_tab;
    // segment mode (kzg_decompress_segment_host): how the next batch's only frame starts
    bool seg_on = false;
    u32 seg_init_owner[4] = { 0, 0, 0, 0 }, seg_block_max = 0;
    DevBuf seg_xxh;
    std::vector<u8> seg_frame;
    // raw-content dictionary applied to every frame of subsequent batches
    DevBuf dict;                   // dictionary content (history prefix)
    u32 dict_size = 0;
    std::vector<u8> dict_host;     // the bytes last loaded (to skip reloading the same dictionary)
    // formatted dictionary (magic 0xEC30A437): entropy tables live behind the constants, see KzgConsts
    u32 dict_id = 0, dict_entropy = 0, dict_rep[3] = { 1, 4, 8 };
    // optional per-stage CUDA events (kzg_ctx_set_profiling)
    // (a ring of event sets: batches enqueued back to back without a host sync in between each keep their own times)
    bool profiling = false;
    cudaEvent_t ev[KZG_EV_RING][KZG_N_STAGES + 1] = {};
    u64 ev_seq = 0;                // profiled batches enqueued so far; the newest uses set (ev_seq - 1) % KZG_EV_RING
    bool ev_valid = false;
};

static const char* err_name(u32 c) {
    switch (c) {
        case 0: return "No error detected";
        case 1: return "Error (generic)";
        case 10: return "Unknown frame descriptor";
        case 12: return "Version not supported";
        case 14: return "Unsupported frame parameter";
        case 16: return "Frame requires too much memory for decoding";
        case 20: return "Data corruption detected";
        case 22: return "Restored data doesn't match checksum";
        case 24: return "Header of Literals' block doesn't respect format specification";
        case 30: return "Dictionary is corrupted";
        case 32: return "Dictionary mismatch";
        case 34: return "Cannot create Dictionary from provided samples";
        case 40: return "Unsupported parameter";
        case 41: return "Unsupported combination of parameters";
        case 42: return "Parameter is out of bound";
        case 44: return "tableLog requires too much memory : unsupported";
        case 46: return "Unsupported max Symbol Value : too large";
        case 48: return "Specified maxSymbolValue is too small";
        case 50: return "pledged buffer stability condition is not respected";
        case 60: return "Operation not authorized at current processing stage";
        case 62: return "Context should be init first";
        case 64: return "Allocation error : not enough memory";
        case 66: return "workSpace buffer is not large enough";
        case 70: return "Destination buffer is too small";
        case 72: return "Src size is incorrect";
        case 74: return "Operation on NULL destination buffer";
        case 80: return "Operation made no progress over multiple calls, due to output buffer being full";
        case 82: return "Operation made no progress over multiple calls, due to input being empty";
        case 100: return "Frame index is too large";
        case 102: return "An I/O error occurred when reading/seeking";
        case 104: return "Destination buffer is wrong";
        case 105: return "Source buffer is wrong";
        case 106: return "Block-level external sequence producer returned an error code";
        case 107: return "External sequences are not valid";
        default: return "Unspecified error code";
    }
}

extern "C" unsigned kzg_isError(size_t code) { return code > (size_t)0 - (size_t)120; }
extern "C" const char* kzg_getErrorName(size_t code) { return kzg_isError(code) ? err_name((u32)((size_t)0 - code)) : err_name(0); }
extern "C" const char* kzg_versionString(void) { return KZG_VERSION_STRING; }

extern "C" int kzg_ctx_create(kzg_ctx** out, int device) {
    if (!out) return (int)KZG_E_GENERIC;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        fprintf(stderr, "kzg: no CUDA device available; this library has no CPU fallback\n");
        return (int)KZG_E_MEMORY;
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) return (int)KZG_E_MEMORY; }
    if (device >= ndev) return (int)KZG_E_MEMORY;
    if (cudaSetDevice(device) != cudaSuccess) return (int)KZG_E_MEMORY;
    kzg_ctx* c = new (std::nothrow) kzg_ctx();
    if (!c) return (int)KZG_E_MEMORY;
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return (int)KZG_E_MEMORY; }
    c->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) fprintf(stderr, "kzg: warning: built for sm_100a, device is sm_%d%d\n", prop.major, prop.minor);
    if (kzg_kernels_init_device(device) != 0) { delete c; return (int)KZG_E_MEMORY; }
    if (cudaMallocHost((void**)&c->h_ctr, KZG_MAX_CHUNKS * sizeof(KzgCounters)) != cudaSuccess) { delete c; return (int)KZG_E_MEMORY; }
    memset(c->h_ctr, 0, KZG_MAX_CHUNKS * sizeof(KzgCounters));
    bool ok = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->in_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_tab, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < KZG_MAX_CHUNKS; i++)
        ok = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_cmp[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { cudaGetLastError(); kzg_ctx_destroy(c); return (int)KZG_E_MEMORY; }
    // constants: default FSE descriptor + packed code tables
    KzgConsts hc;
    memset(&hc, 0, sizeof hc);
    kzg_default_fse_desc(hc.def);
    kzg_code_tables(hc.codes, hc.codes + 36);
    if (c->consts.ensure(sizeof hc) || c->ctr.ensure(sizeof(KzgCounters)) ||
        cudaMemcpy(c->consts.p, &hc, sizeof hc, cudaMemcpyHostToDevice) != cudaSuccess) {
        kzg_ctx_destroy(c);
        return (int)KZG_E_MEMORY;
    }
    *out = c;
    return 0;
}

extern "C" void kzg_ctx_destroy(kzg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    DevBuf* all[] = { &c->frames, &c->blocks, &c->cblocks, &c->fse, &c->huf, &c->hlist, &c->hlist2, &c->slist, &c->blist, &c->dlist, &c->lit, &c->seq, &c->fsetab,
                      &c->ctr, &c->consts, &c->st_src, &c->st_dst, &c->st_tab, &c->dict };
    for (DevBuf* b : all) b->release();
    if (c->h_ctr) cudaFreeHost(c->h_ctr);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->in_stream) cudaStreamDestroy(c->in_stream);
    if (c->out_stream) cudaStreamDestroy(c->out_stream);
    if (c->ev_tab) cudaEventDestroy(c->ev_tab);
    for (int i = 0; i < KZG_MAX_CHUNKS; i++) { if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]); if (c->ev_cmp[i]) cudaEventDestroy(c->ev_cmp[i]); }
    for (int r = 0; r < KZG_EV_RING; r++)
        for (int i = 0; i <= KZG_N_STAGES; i++) if (c->ev[r][i]) cudaEventDestroy(c->ev[r][i]);
    delete c;
}

// Sets (dict != NULL) or clears the dictionary used by the following batches (reference:
// ZSTD_DCtx_loadDictionary at Wrapper.cpp:58-73).  A raw-content dictionary is a history prefix.  A formatted
// dictionary (magic 0xEC30A437: dictID, Huffman table, OF / ML / LL FSE tables, three repeat offsets, content;
// libzstd ZSTD_loadDEntropy order) also preloads every frame's entropy tables and repeat offsets.
// Parses a dictionary: a formatted one (magic 0xEC30A437: dictID, Huffman table, OF / ML / LL FSE tables, three repeat
// offsets, content; libzstd ZSTD_loadDEntropy order) fills fd / hd / rep / id and sets *entropy; any other buffer is
// raw content.  Returns the offset of the content inside `dict`, or a negative code.
static long long parse_dictionary(const void* dict, size_t size, KzgFseDesc& fd, KzgHufDesc& hd, u32 rep[3], u32& id, u32& entropy) {
    const u8* p = (const u8*)dict;
    const u8* const end = p + size;
    id = 0; entropy = 0; rep[0] = 1; rep[1] = 4; rep[2] = 8;
    memset(&fd, 0, sizeof fd);
    memset(&hd, 0, sizeof hd);
    if (size >= 8 && kzg_rd32(p) == KZG_DICT_MAGIC) {
        id = kzg_rd32(p + 4);
        p += 8;
        const u32 hsz = kzg_read_huf_weights(p, (u32)(end - p), hd.weight, hd.nsym, hd.maxbits);
        if (hsz == 0) return -(long long)KZG_E_DICT_CORRUPTED;
        p += hsz;
        const u32 order[3] = { 1, 2, 0 }, max_sym[3] = { 35, 31, 52 }, max_log[3] = { 9, 8, 9 };   // stored as OF, ML, LL
        for (u32 k = 0; k < 3; k++) {
            const u32 t = order[k];
            u32 nsym = 0, log = 0;
            const u32 h = kzg_read_ncount(p, (u32)(end - p), max_sym[t], max_log[t], fd.norm[t], nsym, log);
            if (h == 0) return -(long long)KZG_E_DICT_CORRUPTED;
            fd.kind[t] = 2; fd.log[t] = (u8)log; fd.nsym[t] = (u8)nsym;
            p += h;
        }
        if (p + 12 > end) return -(long long)KZG_E_DICT_CORRUPTED;
        for (u32 k = 0; k < 3; k++) {
            rep[k] = kzg_rd32(p + 4 * k);
            if (rep[k] == 0 || rep[k] > (u64)(end - (p + 12))) return -(long long)KZG_E_DICT_CORRUPTED;
        }
        p += 12;
        entropy = 1;
    }
    return (long long)(p - (const u8*)dict);
}

extern "C" int kzg_ctx_set_dictionary(kzg_ctx* c, const void* dict, size_t size) {
    if (!c) return -(int)KZG_E_GENERIC;
    std::lock_guard<std::mutex> lock(c->mu);
    if (!dict || size == 0) { c->dict_size = 0; c->dict_host.clear(); c->dict_id = 0; c->dict_entropy = 0; return 0; }
    if (size > 0x7FFFFFFFu) return -(int)KZG_E_MEMORY;
    if (c->dict_host.size() == size && memcmp(c->dict_host.data(), dict, size) == 0) return 0;
    CK(cudaSetDevice(c->device));
    KzgFseDesc fd;                           // (on the stack: contexts load dictionaries concurrently)
    KzgHufDesc hd;
    u32 id = 0, entropy = 0, rep[3];
    const long long coff = parse_dictionary(dict, size, fd, hd, rep, id, entropy);
    if (coff < 0) return (int)coff;
    if (entropy) {
        CK(cudaMemcpy((u8*)c->consts.p + offsetof(KzgConsts, dict_fse), &fd, sizeof fd, cudaMemcpyHostToDevice));
        CK(cudaMemcpy((u8*)c->consts.p + offsetof(KzgConsts, dict_huf), &hd, sizeof hd, cudaMemcpyHostToDevice));
    }
    const u8* p = (const u8*)dict + coff;
    const size_t content = size - (size_t)coff;
    if (c->dict.ensure(content + 16)) return -(int)KZG_E_MEMORY;
    if (content) CK(cudaMemcpy(c->dict.p, p, content, cudaMemcpyHostToDevice));
    c->dict_host.assign((const u8*)dict, (const u8*)dict + size);
    c->dict_size = (u32)content;
    c->dict_id = id; c->dict_entropy = entropy;
    c->dict_rep[0] = rep[0]; c->dict_rep[1] = rep[1]; c->dict_rep[2] = rep[2];
    return 0;
}

extern "C" void* kzg_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void kzg_host_free(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------------------
// Seek table of the zstd seekable format (contrib/seekable_format): a skippable frame at the end of the stream,
//   0x184D2A5E | frame size | n entries {compressed size, decompressed size [, checksum]} | n | descriptor | 0x8F92EAB1
// The batch frame writer can append one (kzg_write_seek_table); with it the frame table of a stream comes from one
// pass over 8 bytes per frame instead of a walk over every block header.
#define KZG_SEEK_SKIP_MAGIC 0x184D2A5Eu
#define KZG_SEEK_MAGIC 0x8F92EAB1u

extern "C" size_t kzg_seek_table_bound(size_t n_frames) { return 8 + n_frames * 8 + 9; }

extern "C" size_t kzg_write_seek_table(void* dst, size_t cap, const uint64_t* comp_size, const uint64_t* content_size, size_t n) {
    const size_t need = kzg_seek_table_bound(n);
    if (!dst || cap < need || n > 0x7FFFFFFu) return (size_t)0 - (size_t)KZG_E_DSTSIZE_TOO_SMALL;
    u8* p = (u8*)dst;
    auto w32 = [&](u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24); p += 4; };
    w32(KZG_SEEK_SKIP_MAGIC);
    w32((u32)(need - 8));
    for (size_t i = 0; i < n; i++) {
        if (comp_size[i] > 0xFFFFFFFFull || content_size[i] > 0xFFFFFFFFull) return (size_t)0 - (size_t)KZG_E_GENERIC;
        w32((u32)comp_size[i]);
        w32((u32)content_size[i]);
    }
    w32((u32)n);
    *p++ = 0;                                   // descriptor: no per-frame checksums
    w32(KZG_SEEK_MAGIC);
    return need;
}

// Frame table from a trailing seek table. Returns the number of entries written (the table's own skippable frame is
// the last one), or 0 if there is no table, it does not describe this buffer exactly, or it does not fit max_frames.
static size_t scan_by_seek_table(const u8* p, size_t len, uint64_t* off, uint64_t* size, uint64_t* content, size_t max_frames) {
    if (len < 17 || kzg_rd32(p + len - 4) != KZG_SEEK_MAGIC) return 0;
    const u32 desc = p[len - 5], n = kzg_rd32(p + len - 9);
    if (desc & 0x7C) return 0;
    const size_t entry = (desc & 0x80) ? 12 : 8;
    const u64 table = (u64)n * entry + 9;
    if (table + 8 > len || (u64)n + 1 > max_frames) return 0;
    const size_t tpos = len - (size_t)table - 8;
    if (kzg_rd32(p + tpos) != KZG_SEEK_SKIP_MAGIC || kzg_rd32(p + tpos + 4) != table) return 0;
    u64 pos = 0;
    for (u32 i = 0; i < n; i++) {
        const u8* e = p + tpos + 8 + (size_t)i * entry;
        const u64 cs = kzg_rd32(e), ds = kzg_rd32(e + 4);
        if (cs < 4 || pos + cs > tpos) return 0;
        const u32 magic = kzg_rd32(p + pos);
        if (magic != KZG_MAGIC && (magic & 0xFFFFFFF0u) != KZG_SKIP_MAGIC) return 0;
        if (magic == KZG_MAGIC) {
            // the table is only a hint (other decoders ignore it): the frame's own header must carry the same content
            // size, else the block headers are walked as if there were no table
            KzgFrameHeader h;
            if (kzg_parse_frame_header(p + pos, (size_t)cs, h) != KZG_OK || h.fcs != ds) return 0;
        } else if (8ull + kzg_rd32(p + pos + 4) != cs) return 0;
        if (off) off[i] = pos;
        if (size) size[i] = cs;
        if (content) content[i] = magic == KZG_MAGIC ? ds : 0;
        pos += cs;
    }
    if (pos != tpos) return 0;                  // the entries must cover everything in front of the table
    if (off) off[n] = tpos;
    if (size) size[n] = table + 8;
    if (content) content[n] = 0;
    return (size_t)n + 1;
}

extern "C" size_t kzg_scan_frames(const void* buf, size_t len, uint64_t* off, uint64_t* size, uint64_t* content,
                                  size_t max_frames, size_t* consumed) {
    const u8* p = (const u8*)buf;
    size_t pos = 0, n = 0;
    if (consumed) *consumed = 0;
    if (const size_t k = scan_by_seek_table(p, len, off, size, content, max_frames)) {
        if (consumed) *consumed = len;
        return k;
    }
    while (pos < len && n < max_frames) {
        const u8* f = p + pos;
        const size_t rem = len - pos;
        if (rem < 4) break;
        const u32 magic = kzg_rd32(f);
        if ((magic & 0xFFFFFFF0u) == KZG_SKIP_MAGIC) {
            if (rem < 8) break;
            const u64 fs = 8ull + kzg_rd32(f + 4);
            if (fs > rem) break;
            if (off) off[n] = pos;
            if (size) size[n] = fs;
            if (content) content[n] = 0;
            n++; pos += (size_t)fs;
            if (consumed) *consumed = pos;
            continue;
        }
        if (magic != KZG_MAGIC) return (size_t)0 - (size_t)KZG_E_PREFIX_UNKNOWN;
        KzgFrameHeader h;
        const u32 e = kzg_parse_frame_header(f, rem, h);
        if (e == KZG_E_SRCSIZE_WRONG) break;
        if (e) return (size_t)0 - (size_t)e;
        const u64 bmax = h.window_size < KZG_BLOCK_MAX ? h.window_size : KZG_BLOCK_MAX;
        u64 q = h.hdr_size, bound = 0;
        bool complete = false;
        for (;;) {
            if (q + 3 > rem) break;
            const u32 bh = kzg_rd24(f + q);
            const u32 type = (bh >> 1) & 3, bsz = bh >> 3;
            if (type == 3) return (size_t)0 - (size_t)KZG_E_CORRUPTION;
            const u64 adv = type == 1 ? 1 : bsz;
            if (q + 3 + adv > rem) break;
            bound += type == 2 ? bmax : (bsz < bmax ? bsz : bmax);     // (a raw / RLE block larger than the block maximum is corrupt anyway)
            q += 3 + adv;
            if (bh & 1) { complete = true; break; }
        }
        if (!complete) break;
        if (h.has_checksum) { if (q + 4 > rem) break; q += 4; }
        if (off) off[n] = pos;
        if (size) size[n] = q;
        // The Frame_Content_Size field is untrusted: a frame cannot regenerate more than its blocks allow, so the
        // capacity handed out is never larger than that bound (a frame that claims more fails with corruption_detected,
        // as in libzstd, instead of reserving what it claims).
        if (content) content[n] = (h.fcs != ~0ull && h.fcs < bound) ? h.fcs : bound;
        n++; pos += (size_t)q;
        if (consumed) *consumed = pos;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
static int ensure_workspace(kzg_ctx* c, u64 n_frames, u64 total_src, u64 total_dst, u64 max_frame_dst) {
    // wanted sizes first, in locals: the sticky capacities of the context change only when every allocation succeeded
    u64 block_cap = c->block_cap, lit_cap = c->lit_cap, seq_cap = c->seq_cap;
    u32 tile_bytes = c->tile_bytes;
    u64 want_blocks = n_frames + (u64)((double)total_dst / c->block_div) + 64;
    if (want_blocks > 0xFFFFFFF0ull) want_blocks = 0xFFFFFFF0ull;
    if (want_blocks > block_cap) block_cap = want_blocks;
    const u64 want_lit = (u64)((double)total_dst * c->lit_factor) + n_frames * 48 + (1u << 16);
    if (want_lit > lit_cap) lit_cap = want_lit;
    const u64 want_seq = (u64)((double)total_dst / c->seq_div) + n_frames * 8 + (1u << 12);
    if (want_seq > seq_cap) seq_cap = want_seq;
    u64 blk = max_frame_dst < KZG_BLOCK_MAX ? max_frame_dst : KZG_BLOCK_MAX;
    u32 tile = (u32)((blk + 15) & ~15ull) + 64;
    if (tile < 16384 + 64) tile = 16384 + 64;
    if (tile > tile_bytes) tile_bytes = tile;
    (void)total_src;
    const u64 cblock_cap = block_cap;
    if (c->frames.ensure(n_frames * sizeof(KzgFrame)) || c->dlist.ensure(n_frames * 4 + 16) || c->blocks.ensure(block_cap * sizeof(KzgBlock)) ||
        c->cblocks.ensure(cblock_cap * sizeof(KzgCBlock)) || c->fse.ensure(cblock_cap * sizeof(KzgFseDesc)) ||
        c->huf.ensure(cblock_cap * sizeof(KzgHufDesc)) || c->hlist.ensure(cblock_cap * 4) || c->hlist2.ensure(cblock_cap * 4) ||
        c->slist.ensure(cblock_cap * 4) || c->blist.ensure(block_cap * 4) || c->lit.ensure(lit_cap + 64) || c->seq.ensure(seq_cap * 8) ||
        c->fsetab.ensure(cblock_cap * 2560 + 64)) {
        // a request that does not fit must not poison the context: the next (smaller) batch starts from the defaults
        // again, with whatever buffers survived
        c->block_cap = c->cblock_cap = 0; c->lit_cap = 0; c->seq_cap = 0;
        return -1;
    }
    c->block_cap = block_cap; c->cblock_cap = cblock_cap; c->lit_cap = lit_cap; c->seq_cap = seq_cap; c->tile_bytes = tile_bytes;
    return 0;
}

// Enqueues one batch on `stream` (context already locked / device set).  The batch's counters are copied to
// pinned slot `slot` when it finishes.  Returns kernels launched or a negative code.
static int enqueue_batch(kzg_ctx* c, const void* d_src, const uint64_t* d_src_off, const uint64_t* d_src_size, void* d_dst,
                         const uint64_t* d_dst_off, const uint64_t* d_dst_cap, uint64_t n_frames, uint64_t total_src_bytes,
                         uint64_t total_dst_cap, uint64_t max_frame_dst_cap, uint32_t* d_status, uint64_t* d_out_size,
                         cudaStream_t stream, int slot) {
    if (ensure_workspace(c, n_frames, total_src_bytes, total_dst_cap, max_frame_dst_cap)) return -(int)KZG_E_MEMORY;
    KzgBatch b;
    memset(&b, 0, sizeof b);
    b.src = (const u8*)d_src; b.src_off = d_src_off; b.src_size = d_src_size;
    b.dst = (u8*)d_dst; b.dst_off = d_dst_off; b.dst_cap = d_dst_cap;
    b.status = d_status; b.out_size = d_out_size;
    b.n_frames = (u32)n_frames;
    b.tile_bytes = c->tile_bytes;
    b.frames = (KzgFrame*)c->frames.p; b.blocks = (KzgBlock*)c->blocks.p; b.cblocks = (KzgCBlock*)c->cblocks.p;
    b.fse = (KzgFseDesc*)c->fse.p; b.huf = (KzgHufDesc*)c->huf.p;
    b.hlist = (u32*)c->hlist.p; b.hlist2 = (u32*)c->hlist2.p; b.slist = (u32*)c->slist.p; b.blist = (u32*)c->blist.p; b.dlist = (u32*)c->dlist.p;
    b.lit_pool = (u8*)c->lit.p; b.seq_pool = (u64*)c->seq.p; b.fse_tab = (u16*)c->fsetab.p;
    b.ctr = (KzgCounters*)c->ctr.p;
    b.fse_default = (const KzgFseDesc*)((const u8*)c->consts.p + offsetof(KzgConsts, def));
    b.code_tables = (const u32*)((const u8*)c->consts.p + offsetof(KzgConsts, codes));
    b.fse_dict = (const KzgFseDesc*)((const u8*)c->consts.p + offsetof(KzgConsts, dict_fse));
    b.huf_dict = (const KzgHufDesc*)((const u8*)c->consts.p + offsetof(KzgConsts, dict_huf));
    b.lit_cap = c->lit_cap; b.seq_cap = c->seq_cap;
    b.block_cap = (u32)c->block_cap; b.cblock_cap = (u32)c->cblock_cap;
    b.dict = (const u8*)c->dict.p; b.dict_size = c->dict_size; b.dict_id = c->dict_id;
    b.dict_entropy = c->dict_entropy;
    b.dict_rep[0] = c->dict_rep[0]; b.dict_rep[1] = c->dict_rep[1]; b.dict_rep[2] = c->dict_rep[2];
    if (c->seg_on) {
        b.seg_mode = 1; b.seg_block_max = c->seg_block_max;
        for (int t = 0; t < 4; t++) b.init_owner[t] = c->seg_init_owner[t];
    }
    CK(cudaMemsetAsync(c->ctr.p, 0, sizeof(KzgCounters), stream));
    const int launches = kzg_launch_batch(b, c->sm_count, stream, c->profiling ? c->ev[c->ev_seq % KZG_EV_RING] : nullptr);
    if (c->profiling) c->ev_seq++;
    c->ev_valid = c->profiling;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(c->h_ctr + slot, c->ctr.p, sizeof(KzgCounters), cudaMemcpyDeviceToHost, stream));
    return launches;
}

extern "C" int kzg_decompress_batch_device(kzg_ctx* c, const void* d_src, const uint64_t* d_src_off,
                                           const uint64_t* d_src_size, void* d_dst, const uint64_t* d_dst_off,
                                           const uint64_t* d_dst_cap, uint64_t n_frames, uint64_t total_src_bytes,
                                           uint64_t total_dst_cap, uint64_t max_frame_dst_cap, uint32_t* d_status,
                                           uint64_t* d_out_size, void* cuda_stream) {
    if (!c) return -(int)KZG_E_GENERIC;
    if (n_frames > 0xFFFFFFF0ull) return -(int)KZG_E_GENERIC;
    std::lock_guard<std::mutex> lock(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    c->last_stream = stream;
    c->last_n_frames = n_frames;
    c->last_launches = 0;
    if (n_frames == 0) return 0;
    const int launches = enqueue_batch(c, d_src, d_src_off, d_src_size, d_dst, d_dst_off, d_dst_cap, n_frames, total_src_bytes,
                                       total_dst_cap, max_frame_dst_cap, d_status, d_out_size, stream, 0);
    if (launches > 0) c->last_launches = (u32)launches;
    return launches;
}

// Grows the workspace after an overflow reported in counters `k`.
static void grow_after_overflow(kzg_ctx* c, const KzgCounters& k) {
    if (k.overflow & 3u) {
        const u64 need = (u64)(k.n_blocks > k.n_cblocks ? k.n_blocks : k.n_cblocks) + 64;
        c->block_cap = c->cblock_cap = need > c->block_cap * 2 ? need : c->block_cap * 2;
    }
    if (k.overflow & 4u) { const u64 w = k.lit_used + (k.lit_used >> 4) + 4096; if (w > c->lit_cap) c->lit_cap = w; }
    if (k.overflow & 8u) { const u64 w = k.seq_used + (k.seq_used >> 4) + 4096; if (w > c->seq_cap) c->seq_cap = w; }
    if (k.overflow & 16u) c->tile_bytes = KZG_BLOCK_MAX + 64;
}

extern "C" int kzg_batch_finish(kzg_ctx* c) {
    if (!c) return -(int)KZG_E_GENERIC;
    std::lock_guard<std::mutex> lock(c->mu);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->last_stream));
    if (c->last_n_frames == 0) return 0;
    const KzgCounters& k = c->h_ctr[0];
    if (!k.overflow) return 0;
    grow_after_overflow(c, k);
    return 1;
}

extern "C" int kzg_ctx_set_profiling(kzg_ctx* c, int enable) {
    if (!c) return -(int)KZG_E_GENERIC;
    std::lock_guard<std::mutex> lock(c->mu);
    CK(cudaSetDevice(c->device));
    if (enable && !c->ev[0][0])
        for (int r = 0; r < KZG_EV_RING; r++)
            for (int i = 0; i <= KZG_N_STAGES; i++) CK(cudaEventCreate(&c->ev[r][i]));
    c->profiling = enable != 0;
    c->ev_valid = false;
    c->ev_seq = 0;
    return 0;
}

extern "C" int kzg_stage_ms_history(kzg_ctx* c, int age, float* ms, const char** names, int max_stages) {
    if (!c || !ms || age < 0) return -(int)KZG_E_GENERIC;
    if (!c->ev_valid || (u64)age >= c->ev_seq || age >= KZG_EV_RING) return 0;
    CK(cudaSetDevice(c->device));
    cudaEvent_t* ev = c->ev[(c->ev_seq - 1 - (u64)age) % KZG_EV_RING];
    int n = max_stages < KZG_N_STAGES ? max_stages : KZG_N_STAGES;
    for (int i = 0; i < n; i++) {
        CK(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
        if (names) names[i] = kzg_stage_names[i];
    }
    return n;
}

extern "C" int kzg_last_stage_ms(kzg_ctx* c, float* ms, const char** names, int max_stages) {
    return kzg_stage_ms_history(c, 0, ms, names, max_stages);
}

extern "C" int kzg_last_batch_stats(kzg_ctx* c, kzg_batch_stats* out) {
    if (!c || !out) return -(int)KZG_E_GENERIC;
    const KzgCounters& k = *c->h_ctr;
    out->n_blocks = k.n_blocks; out->n_cblocks = k.n_cblocks; out->n_hblocks = k.n_hblocks; out->n_sblocks = k.n_sblocks;
    out->lit_bytes = k.lit_used; out->seq_records = k.seq_used;
    out->workspace_bytes = c->frames.cap + c->blocks.cap + c->cblocks.cap + c->fse.cap + c->huf.cap + c->hlist.cap +
                           c->slist.cap + c->lit.cap + c->seq.cap + c->fsetab.cap;
    out->launches = c->last_launches; out->tile_bytes = c->tile_bytes; out->sm_count = (u32)c->sm_count;
    out->overflow = k.overflow;
    return 0;
}

// ------------------------------------------------------------------------------------------------
static int batch_host_locked(kzg_ctx* c, const void* src, const uint64_t* src_off, const uint64_t* src_size,
                             void* dst, const uint64_t* dst_off, const uint64_t* dst_cap, uint64_t n,
                             uint32_t* status, uint64_t* out_size);

extern "C" int kzg_decompress_batch_host(kzg_ctx* c, const void* src, const uint64_t* src_off, const uint64_t* src_size,
                                         void* dst, const uint64_t* dst_off, const uint64_t* dst_cap, uint64_t n,
                                         uint32_t* status, uint64_t* out_size) {
    if (!c) return -(int)KZG_E_GENERIC;
    if (n == 0) return 0;
    if (n > 0xFFFFFFF0ull) return -(int)KZG_E_GENERIC;
    std::lock_guard<std::mutex> lock(c->mu);
    return batch_host_locked(c, src, src_off, src_size, dst, dst_off, dst_cap, n, status, out_size);
}

static int batch_host_locked(kzg_ctx* c, const void* src, const uint64_t* src_off, const uint64_t* src_size,
                             void* dst, const uint64_t* dst_off, const uint64_t* dst_cap, uint64_t n,
                             uint32_t* status, uint64_t* out_size) {
    CK(cudaSetDevice(c->device));
    // extent of the source / destination spans
    u64 src_lo = ~0ull, src_hi = 0, dst_lo = ~0ull, dst_hi = 0, tot_dst = 0;
    for (u64 i = 0; i < n; i++) {
        // (sizes come from untrusted frame headers further up the stack: no sum may wrap)
        const u64 se = src_off[i] + src_size[i], de = dst_off[i] + dst_cap[i];
        if (se < src_off[i] || de < dst_off[i] || tot_dst + dst_cap[i] < tot_dst) return -(int)KZG_E_GENERIC;
        if (src_off[i] < src_lo) src_lo = src_off[i];
        if (se > src_hi) src_hi = se;
        if (dst_off[i] < dst_lo) dst_lo = dst_off[i];
        if (de > dst_hi) dst_hi = de;
        tot_dst += dst_cap[i];
    }
    if (src_hi - src_lo > (1ull << 46) || dst_hi - dst_lo > (1ull << 46)) return -(int)KZG_E_MEMORY;     // far beyond any device
    const u64 src_span = src_hi - src_lo, dst_span = dst_hi - dst_lo;
    // device staging: [src span | pad], [dst span], tables: off/size/doff/dcap (u64 each) + out_size (u64) + status (u32)
    const u64 tab_bytes = n * (5 * 8 + 4);
    if (c->st_src.ensure(src_span + 64) || c->st_dst.ensure(dst_span + 64) || c->st_tab.ensure(tab_bytes + 64))
        return -(int)KZG_E_MEMORY;
    u64* t_off = (u64*)c->st_tab.p;
    u64* t_size = t_off + n;
    u64* t_doff = t_size + n;
    u64* t_dcap = t_doff + n;
    u64* t_osize = t_dcap + n;
    u32* t_status = (u32*)(t_osize + n);
    // rebased offset tables
    u64* h_tab = (u64*)malloc(n * 4 * 8);
    if (!h_tab) return -(int)KZG_E_MEMORY;
    for (u64 i = 0; i < n; i++) {
        h_tab[i] = src_off[i] - src_lo; h_tab[n + i] = src_size[i];
        h_tab[2 * n + i] = dst_off[i] - dst_lo; h_tab[3 * n + i] = dst_cap[i];
    }
    // Sub-batches of consecutive frames, pipelined over three streams: while sub-batch k is decoded, the
    // compressed bytes of k+1 are on their way in and the output of k-1 is on its way out (the two PCIe
    // directions and the SMs all busy at once).  The kernels of successive sub-batches run back to back on one
    // stream, so they can share the workspace.
    // The first sub-batch is half the target size (then the target) so that the first output starts
    // crossing PCIe early; the target is large enough for a sub-batch to still fill the GPU.
    // (KZG_CHUNK_MB / KZG_CHUNK_RAMP: tuning knobs read once from the environment.)
    static const u64 chunk_bytes = [] { const char* e = getenv("KZG_CHUNK_MB"); const u64 v = e ? strtoull(e, nullptr, 10) : 0; return v ? v << 20 : KZG_CHUNK_BYTES; }();
    static const u32 ramp_shift = [] { const char* e = getenv("KZG_CHUNK_RAMP"); return e ? (u32)strtoul(e, nullptr, 10) & 7u : 1u; }();
    u64 nchunks = 0;
    u64 bounds[KZG_MAX_CHUNKS + 1];
    bounds[0] = 0;
    if (tot_dst <= chunk_bytes) { nchunks = 1; bounds[1] = n; }
    else {
        u64 acc = 0, goal = chunk_bytes >> ramp_shift, i = 0;
        while (i < n) {
            acc += dst_cap[i++];
            if ((acc >= goal && nchunks + 1 < KZG_MAX_CHUNKS) || i == n) {
                bounds[++nchunks] = i;
                acc = 0;
                if (goal < chunk_bytes) goal <<= 1;
                else if (nchunks + 4 >= KZG_MAX_CHUNKS) goal <<= 1;      // very large calls: keep within the chunk limit
            }
        }
    }
    cudaStream_t s_in = c->in_stream, s_cmp = c->own_stream, s_out = c->out_stream;
    c->last_stream = s_cmp;
    c->last_n_frames = n;
    int rc = 0;
    u32 launches_total = 0;
    for (int attempt = 0; attempt < 4; attempt++) {
        bool ok = cudaMemcpyAsync(c->st_tab.p, h_tab, n * 4 * 8, cudaMemcpyHostToDevice, s_in) == cudaSuccess &&
                  cudaEventRecord(c->ev_tab, s_in) == cudaSuccess && cudaStreamWaitEvent(s_cmp, c->ev_tab, 0) == cudaSuccess;
        launches_total = 0;
        rc = 0;
        for (u64 k = 0; ok && k < nchunks; k++) {
            const u64 f0 = bounds[k], f1 = bounds[k + 1];
            if (f1 <= f0) { cudaEventRecord(c->ev_in[k], s_in); cudaEventRecord(c->ev_cmp[k], s_cmp); continue; }
            // compressed bytes of this sub-batch: the span its frames cover
            u64 lo = ~0ull, hi = 0, cs = 0, cd = 0, md = 0;
            for (u64 i = f0; i < f1; i++) {
                if (src_off[i] < lo) lo = src_off[i];
                if (src_off[i] + src_size[i] > hi) hi = src_off[i] + src_size[i];
                cs += src_size[i]; cd += dst_cap[i];
                if (dst_cap[i] > md) md = dst_cap[i];
            }
            ok = cudaMemcpyAsync((u8*)c->st_src.p + (lo - src_lo), (const u8*)src + lo, hi - lo, cudaMemcpyHostToDevice, s_in) == cudaSuccess &&
                 cudaEventRecord(c->ev_in[k], s_in) == cudaSuccess && cudaStreamWaitEvent(s_cmp, c->ev_in[k], 0) == cudaSuccess;
            if (!ok) break;
            const int l = enqueue_batch(c, c->st_src.p, t_off + f0, t_size + f0, c->st_dst.p, t_doff + f0, t_dcap + f0, f1 - f0, cs, cd, md,
                                        t_status + f0, t_osize + f0, s_cmp, (int)k);
            if (l < 0) { rc = l; break; }
            launches_total += (u32)l;
            ok = cudaEventRecord(c->ev_cmp[k], s_cmp) == cudaSuccess && cudaStreamWaitEvent(s_out, c->ev_cmp[k], 0) == cudaSuccess;
            // copy back only the caller's capacity regions of this sub-batch, coalescing adjacent ones into runs
            u64 run_lo = dst_off[f0], run_hi = dst_off[f0] + dst_cap[f0];
            for (u64 i = f0 + 1; ok && i <= f1; i++) {
                if (i < f1 && dst_off[i] == run_hi) { run_hi += dst_cap[i]; continue; }
                if (run_hi > run_lo)
                    ok = cudaMemcpyAsync((u8*)dst + run_lo, (const u8*)c->st_dst.p + (run_lo - dst_lo), run_hi - run_lo,
                                         cudaMemcpyDeviceToHost, s_out) == cudaSuccess;
                if (i < f1) { run_lo = dst_off[i]; run_hi = dst_off[i] + dst_cap[i]; }
            }
        }
        if (rc == 0 && ok)
            ok = cudaMemcpyAsync(status, t_status, n * 4, cudaMemcpyDeviceToHost, s_out) == cudaSuccess &&
                 cudaMemcpyAsync(out_size, t_osize, n * 8, cudaMemcpyDeviceToHost, s_out) == cudaSuccess;
        const bool synced = cudaStreamSynchronize(s_in) == cudaSuccess && cudaStreamSynchronize(s_cmp) == cudaSuccess &&
                            cudaStreamSynchronize(s_out) == cudaSuccess;
        if (rc < 0) break;
        if (!ok || !synced) { cudaGetLastError(); rc = -(int)KZG_E_MEMORY; break; }
        // optimistic workspace sizing: any sub-batch that overflowed makes the whole call run again, larger
        bool again = false;
        for (u64 k = 0; k < nchunks; k++)
            if (bounds[k + 1] > bounds[k] && c->h_ctr[k].overflow) { grow_after_overflow(c, c->h_ctr[k]); again = true; }
        if (!again) break;
        rc = 1;
    }
    if (rc == 1) rc = -(int)KZG_E_MEMORY;
    c->last_launches = launches_total;
    free(h_tab);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// A frame decoded in segments (include/kzg.h).  A segment runs through the very same batch pipeline as a synthetic
// frame -- magic, a header without content size / checksum / dictionary ID, the blocks -- in "segment mode": the tables
// that repeat / treeless modes refer to, the repeat offsets and the history come from the state instead of a frame
// start.  Afterwards the state is advanced from what the link and offsets stages left in the frame descriptor.
static_assert(sizeof(((kzg_seg_state*)0)->fse_desc) == sizeof(KzgFseDesc), "kzg_seg_state.fse_desc holds a KzgFseDesc");
static_assert(sizeof(((kzg_seg_state*)0)->huf_desc) == sizeof(KzgHufDesc), "kzg_seg_state.huf_desc holds a KzgHufDesc");

extern "C" long long kzg_segment_begin(kzg_seg_state* st, uint32_t block_max, int want_checksum, const void* dict, size_t dict_size) {
    if (!st) return -(long long)KZG_E_GENERIC;
    memset(st, 0, sizeof *st);
    st->rep[0] = 1; st->rep[1] = 4; st->rep[2] = 8;
    st->block_max = block_max;
    st->want_checksum = want_checksum ? 1u : 0u;
    st->xxh_acc[0] = 0x9E3779B185EBCA87ull + 0xC2B2AE3D27D4EB4Full; st->xxh_acc[1] = 0xC2B2AE3D27D4EB4Full;
    st->xxh_acc[2] = 0; st->xxh_acc[3] = 0ull - 0x9E3779B185EBCA87ull;
    st->digest = 0x51D8E999u;                 // low half of XXH64 of the empty input
    if (!dict || dict_size == 0) return 0;
    KzgFseDesc fd;
    KzgHufDesc hd;
    u32 id = 0, entropy = 0, rep[3];
    const long long coff = parse_dictionary(dict, dict_size, fd, hd, rep, id, entropy);
    if (coff < 0) return coff;
    if (entropy) {
        memcpy(st->fse_desc, &fd, sizeof fd);
        memcpy(st->huf_desc, &hd, sizeof hd);
        st->owner_kind[0] = st->owner_kind[1] = st->owner_kind[2] = st->owner_kind[3] = 2;
        st->rep[0] = rep[0]; st->rep[1] = rep[1]; st->rep[2] = rep[2];
    }
    return (long long)dict_size - coff;
}

extern "C" uint32_t kzg_segment_checksum(const kzg_seg_state* st) { return st ? st->digest : 0u; }

extern "C" int kzg_decompress_segment_host(kzg_ctx* c, kzg_seg_state* st, const void* blocks, size_t blocks_size, size_t last_hdr_off,
                                           const void* history, size_t history_size, void* dst, size_t dst_cap,
                                           uint32_t* status, uint64_t* out_size) {
    if (!c || !st || !blocks || !status || !out_size || blocks_size < 3 || last_hdr_off + 3 > blocks_size) return -(int)KZG_E_GENERIC;
    if (history_size > 0x7FFFFFFFu || blocks_size > 0x7FFFFFF0u) return -(int)KZG_E_MEMORY;
    std::lock_guard<std::mutex> lock(c->mu);
    CK(cudaSetDevice(c->device));
    // the synthetic frame: magic | descriptor 0 (window descriptor follows; no content size, checksum, dictionary ID) |
    // window descriptor (the block size limit comes from the state) | blocks, the last one marked last
    try { c->seg_frame.resize(6 + blocks_size); } catch (const std::bad_alloc&) { return -(int)KZG_E_MEMORY; }
    u8* f = c->seg_frame.data();
    f[0] = 0x28; f[1] = 0xB5; f[2] = 0x2F; f[3] = 0xFD; f[4] = 0x00; f[5] = 0x00;
    memcpy(f + 6, blocks, blocks_size);
    f[6 + last_hdr_off] |= 1u;
    // history and tables
    if (c->dict.ensure(history_size + 16)) return -(int)KZG_E_MEMORY;
    if (history_size) CK(cudaMemcpy(c->dict.p, history, history_size, cudaMemcpyHostToDevice));
    CK(cudaMemcpy((u8*)c->consts.p + offsetof(KzgConsts, dict_fse), st->fse_desc, sizeof(KzgFseDesc), cudaMemcpyHostToDevice));
    CK(cudaMemcpy((u8*)c->consts.p + offsetof(KzgConsts, dict_huf), st->huf_desc, sizeof(KzgHufDesc), cudaMemcpyHostToDevice));
    c->dict_host.clear();                     // (whatever dictionary was loaded is gone: the next kzg_ctx_set_dictionary reloads)
    c->dict_size = (u32)history_size; c->dict_id = 0; c->dict_entropy = 1;
    c->dict_rep[0] = st->rep[0]; c->dict_rep[1] = st->rep[1]; c->dict_rep[2] = st->rep[2];
    c->seg_on = true; c->seg_block_max = st->block_max;
    for (int t = 0; t < 4; t++)
        c->seg_init_owner[t] = st->owner_kind[t] == 2 ? KZG_DICT_OWNER : (st->owner_kind[t] == 1 && t ? KZG_NONE : KZG_INVALID_OWNER);
    const uint64_t off0 = 0, size0 = 6 + blocks_size, cap0 = dst_cap;
    int rc = batch_host_locked(c, f, &off0, &size0, dst, &off0, &cap0, 1, status, out_size);
    c->seg_on = false;
    c->dict_size = 0; c->dict_entropy = 0; c->dict_rep[0] = 1; c->dict_rep[1] = 4; c->dict_rep[2] = 8;
    if (rc < 0) return rc;
    if (*status != 0) return 0;
    // ---- advance the state: table owners and repeat offsets the link / offsets stages left in the frame descriptor
    KzgFrame fr;
    CK(cudaMemcpy(&fr, c->frames.p, sizeof fr, cudaMemcpyDeviceToHost));
    KzgFseDesc fd;
    memcpy(&fd, st->fse_desc, sizeof fd);
    for (int t = 0; t < 4; t++) {
        const u32 o = fr.end_owner[t];
        if (o == KZG_DICT_OWNER) continue;                       // unchanged by this segment
        if (o == KZG_INVALID_OWNER) { st->owner_kind[t] = 0; continue; }
        if (o == KZG_NONE) { st->owner_kind[t] = 1; continue; }
        if (t == 0) CK(cudaMemcpy(st->huf_desc, (const u8*)c->huf.p + (size_t)o * sizeof(KzgHufDesc), sizeof(KzgHufDesc), cudaMemcpyDeviceToHost));
        else {
            KzgFseDesc src;
            CK(cudaMemcpy(&src, (const u8*)c->fse.p + (size_t)o * sizeof(KzgFseDesc), sizeof src, cudaMemcpyDeviceToHost));
            const int k = t - 1;
            memcpy(fd.norm[k], src.norm[k], sizeof fd.norm[k]);
            fd.log[k] = src.log[k]; fd.nsym[k] = src.nsym[k]; fd.kind[k] = src.kind[k];
        }
        st->owner_kind[t] = 2;
    }
    memcpy(st->fse_desc, &fd, sizeof fd);
    st->rep[0] = fr.end_rep[0]; st->rep[1] = fr.end_rep[1]; st->rep[2] = fr.end_rep[2];
    // ---- the running checksum: the segment's output is still in the device staging buffer
    if (st->want_checksum) {
        if (c->seg_xxh.ensure(sizeof(KzgXxhState))) return -(int)KZG_E_MEMORY;
        KzgXxhState xs;
        memcpy(xs.acc, st->xxh_acc, sizeof xs.acc);
        xs.total = st->xxh_total;
        memcpy(xs.tail, st->xxh_tail, 32);
        xs.want = 1; xs.digest = st->digest;
        CK(cudaMemcpyAsync(c->seg_xxh.p, &xs, sizeof xs, cudaMemcpyHostToDevice, c->own_stream));
        kzg_launch_xxh_segment((KzgXxhState*)c->seg_xxh.p, (const u8*)c->st_dst.p, *out_size, c->own_stream);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&xs, c->seg_xxh.p, sizeof xs, cudaMemcpyDeviceToHost, c->own_stream));
        CK(cudaStreamSynchronize(c->own_stream));
        memcpy(st->xxh_acc, xs.acc, sizeof xs.acc);
        st->xxh_total = xs.total;
        memcpy(st->xxh_tail, xs.tail, 32);
        st->digest = xs.digest;
    }
    return 0;
}
