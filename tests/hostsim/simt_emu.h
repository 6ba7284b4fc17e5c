// TEST INFRASTRUCTURE ONLY.  Lock-step warp emulator for the CPU: the host side of kompressor_b200/csrc/kzg_simt.cuh.
//
// The 32 lanes of a warp are 32 coroutines (ucontext).  A lane runs until it reaches a collective (shuffle, ballot,
// any, syncwarp); the scheduler resumes the lanes once all 32 have arrived at the SAME call site, so divergent
// collectives, early exits and (because lanes between two collectives run strictly one after another, in forward or
// reverse lane order) missing barriers around shared-memory exchanges show up as wrong output or as an abort.
// Asynchronous copies are completed as LATE as the program allows: a cp.async lands when its issuing lane's
// wait_group lets it, a bulk shared -> global copy reads shared memory and writes global memory only at the wait.
// Never linked into libkzg.so.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <vector>

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{ x, y, z, w }; }

namespace simt {

enum Op { OP_NONE = 0, OP_SHFL, OP_SHFL_UP, OP_BALLOT, OP_ANY, OP_SYNC, OP_OR, OP_DONE };

struct PendingCp { uint32_t group; uint32_t dst; const uint8_t* src; };
struct PendingBulk { uint32_t group; uint8_t* g; uint32_t s; uint32_t n; std::vector<uint8_t> snap; };

struct Warp {
    static constexpr int W = 32;
    static constexpr uint32_t SMEM_BASE = 0x1000;       // shared "addresses" are offsets from here: 0 stays invalid
    ucontext_t sched;
    ucontext_t ctx[W];
    std::vector<char> stack[W];
    int cur = -1;
    bool reverse = false;
    Op op[W];
    int site[W];
    uint32_t a0[W], a1[W], res[W];
    std::vector<uint8_t> smem;
    std::vector<PendingCp> cp[W];
    uint32_t cp_group[W];
    std::vector<PendingBulk> bulk[W];
    uint32_t bulk_group[W];
    void (*fn)(void*, uint32_t) = nullptr;
    void* arg = nullptr;
    uint64_t n_collectives = 0;
};

static thread_local Warp* g_w = nullptr;
static bool g_cp_eager = false;        // cp.async copies land at once instead of as late as their wait allows

[[noreturn]] static inline void die(const char* msg, int site = 0) {
    fprintf(stderr, "simt_emu: %s (site %d, lane %d)\n", msg, site, g_w ? g_w->cur : -1);
    abort();
}

static inline uint32_t collective(Op op, int site, uint32_t a0, uint32_t a1) {
    Warp* w = g_w;
    const int l = w->cur;
    w->op[l] = op; w->site[l] = site; w->a0[l] = a0; w->a1[l] = a1;
    swapcontext(&w->ctx[l], &w->sched);
    return w->res[l];
}

static void lane_entry(int lane) {
    Warp* w = g_w;
    w->fn(w->arg, (uint32_t)lane);
    w->op[lane] = OP_DONE; w->site[lane] = 0;
    swapcontext(&w->ctx[lane], &w->sched);
    die("resumed a finished lane");
}

// Runs fn(arg, lane) for the 32 lanes of one warp with `smem_bytes` of shared memory.
static inline void run_warp(void (*fn)(void*, uint32_t), void* arg, size_t smem_bytes, bool reverse) {
    Warp w;
    w.fn = fn; w.arg = arg; w.reverse = reverse;
    w.smem.assign(smem_bytes + 64, 0xA5);
    g_w = &w;
    for (int l = 0; l < Warp::W; l++) {
        w.stack[l].resize(256 * 1024);
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stack[l].data();
        w.ctx[l].uc_stack.ss_size = w.stack[l].size();
        w.ctx[l].uc_link = &w.sched;
        makecontext(&w.ctx[l], (void (*)())lane_entry, 1, l);
        w.op[l] = OP_NONE; w.cp_group[l] = 0; w.bulk_group[l] = 0;
    }
    for (;;) {
        for (int k = 0; k < Warp::W; k++) {
            const int l = reverse ? Warp::W - 1 - k : k;
            w.cur = l;
            swapcontext(&w.sched, &w.ctx[l]);
        }
        w.cur = -1;
        const Op op = w.op[0];
        for (int l = 1; l < Warp::W; l++)
            if (w.op[l] != op || w.site[l] != w.site[0]) {
                fprintf(stderr, "simt_emu: divergent collective: lane 0 at op %d site %d, lane %d at op %d site %d\n", (int)op, w.site[0], l, (int)w.op[l], w.site[l]);
                abort();
            }
        if (op == OP_DONE) break;
        w.n_collectives++;
        uint32_t mask = 0, orv = 0;
        for (int l = 0; l < Warp::W; l++) { if (w.a0[l]) mask |= 1u << l; orv |= w.a0[l]; }
        for (int l = 0; l < Warp::W; l++) {
            switch (op) {
                case OP_SHFL: w.res[l] = w.a0[w.a1[l] & 31u]; break;
                case OP_SHFL_UP: w.res[l] = (uint32_t)l >= w.a1[l] ? w.a0[l - w.a1[l]] : w.a0[l]; break;
                case OP_BALLOT: w.res[l] = mask; break;
                case OP_ANY: w.res[l] = mask != 0; break;
                case OP_OR: w.res[l] = orv; break;
                default: w.res[l] = 0; break;
            }
        }
    }
    for (int l = 0; l < Warp::W; l++) {
        if (!w.cp[l].empty()) die("lane exits with cp.async copies that were never waited for");
        if (!w.bulk[l].empty()) die("lane exits with bulk copies that were never waited for");
    }
    g_w = nullptr;
}

#define KZG_SITE __builtin_LINE()
static inline uint32_t w_shfl(uint32_t v, uint32_t src, int site = KZG_SITE) { return collective(OP_SHFL, site, v, src); }
static inline uint32_t w_shfl_up(uint32_t v, uint32_t d, int site = KZG_SITE) { return collective(OP_SHFL_UP, site, v, d); }
static inline uint32_t w_ballot(bool p, int site = KZG_SITE) { return collective(OP_BALLOT, site, p ? 1u : 0u, 0); }
static inline bool w_any(bool p, int site = KZG_SITE) { return collective(OP_ANY, site, p ? 1u : 0u, 0) != 0; }
static inline void w_sync(int site = KZG_SITE) { collective(OP_SYNC, site, 0, 0); }
static inline uint32_t w_or(uint32_t v, int site = KZG_SITE) { return collective(OP_OR, site, v, 0); }
static inline void fence_block() {}
static inline void fence_gpu() {}

static inline uint8_t* sm(uint32_t a, uint32_t n, uint32_t align) {
    Warp* w = g_w;
    if (a < Warp::SMEM_BASE || (size_t)(a - Warp::SMEM_BASE) + n > w->smem.size() - 64) die("shared-memory access out of bounds");
    if (a & (align - 1)) die("misaligned shared-memory access");
    return w->smem.data() + (a - Warp::SMEM_BASE);
}
static inline uint32_t lds32(uint32_t a) { uint32_t v; memcpy(&v, sm(a, 4, 4), 4); return v; }
static inline uint32_t lds8(uint32_t a) { return *sm(a, 1, 1); }
static inline void sts32(uint32_t a, uint32_t v) { memcpy(sm(a, 4, 4), &v, 4); }
static inline void sts8(uint32_t a, uint32_t v) { *sm(a, 1, 1) = (uint8_t)v; }
static inline void sts32_if(bool p, uint32_t a, uint32_t v) { if (p) sts32(a, v); }
static inline void sts8_if(bool p, uint32_t a, uint32_t v) { if (p) sts8(a, v); }
static inline uint4 lds128(uint32_t a) { uint4 v; memcpy(&v, sm(a, 16, 16), 16); return v; }
static inline void sts128(uint32_t a, uint4 v) { memcpy(sm(a, 16, 16), &v, 16); }
static inline void sts64(uint32_t a, uint32_t x, uint32_t y) { uint32_t v[2] = { x, y }; memcpy(sm(a, 8, 8), v, 8); }
static inline uint2 lds64(uint32_t a) { uint2 v; memcpy(&v, sm(a, 8, 8), 8); return v; }

static inline void cp_async16_if(bool p, uint32_t dst, const void* src) {
    if (!p) return;
    Warp* w = g_w;
    sm(dst, 16, 16);
    if ((uintptr_t)src & 15) die("misaligned cp.async source");
    if (g_cp_eager) { memcpy(sm(dst, 16, 16), src, 16); return; }
    w->cp[w->cur].push_back(PendingCp{ w->cp_group[w->cur], dst, (const uint8_t*)src });
}
static inline void cp_commit() { g_w->cp_group[g_w->cur]++; }
template <int N> static inline void cp_wait() {
    Warp* w = g_w;
    const int l = w->cur;
    // groups numbered < cp_group - N are complete
    std::vector<PendingCp> keep;
    for (const PendingCp& c : w->cp[l]) {
        if (c.group + (uint32_t)N < w->cp_group[l]) memcpy(sm(c.dst, 16, 16), c.src, 16);
        else keep.push_back(c);
    }
    w->cp[l].swap(keep);
}

static inline void fence_async() {}
static inline void bulk_s2g(void* g, uint32_t s, uint32_t n) {
    if (((uintptr_t)g & 15) || (s & 15) || (n & 15) || n == 0) die("bulk copy: bad alignment or size");
    sm(s, n, 16);
    g_w->bulk[g_w->cur].push_back(PendingBulk{ g_w->bulk_group[g_w->cur], (uint8_t*)g, s, n, {} });
}
static inline void bulk_commit() { g_w->bulk_group[g_w->cur]++; }
template <int N> static inline void bulk_wait() {        // at most N groups stay pending
    Warp* w = g_w;
    const int l = w->cur;
    std::vector<PendingBulk> keep;
    for (const PendingBulk& b : w->bulk[l]) {
        if (b.group + (uint32_t)N < w->bulk_group[l]) {
            if (b.snap.empty()) memcpy(b.g, sm(b.s, b.n, 16), b.n);
            else memcpy(b.g, b.snap.data(), b.n);
        } else keep.push_back(b);
    }
    w->bulk[l].swap(keep);
}
template <int N> static inline void bulk_wait_read() {   // at most N groups have not read their source yet
    Warp* w = g_w;
    const int l = w->cur;
    for (PendingBulk& b : w->bulk[l])
        if (b.group + (uint32_t)N < w->bulk_group[l] && b.snap.empty()) { const uint8_t* p = sm(b.s, b.n, 16); b.snap.assign(p, p + b.n); }
}
static inline void bulk_wait_all() { bulk_wait<0>(); }

static inline uint32_t ldg32_hist(const uint32_t* p) { return *p; }
static inline uint32_t ldg8_hist(const uint8_t* p) { return *p; }
static inline uint32_t ldg8_ro(const uint8_t* p) { return *p; }
static inline uint32_t ldg32_ro(const uint32_t* p) { return *p; }
static inline uint64_t ldg64_stream(const uint64_t* p) { return *p; }
static inline void stg8(uint8_t* p, uint32_t v) { *p = (uint8_t)v; }
static inline void stg128(uint8_t* p, uint4 v) { if ((uintptr_t)p & 15) die("misaligned 16-byte global store"); memcpy(p, &v, 16); }
static inline uint64_t ldg64_ro(const uint64_t* p) { return *p; }
static inline uint32_t atomic_add(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p += v; return o; }

static inline uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31u; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31u; return sh ? (hi << sh) | (lo >> (32 - sh)) : hi; }
static inline uint32_t shl_clamp(uint32_t v, uint32_t n) { return n >= 32 ? 0u : v << n; }
static inline uint32_t shr_clamp(uint32_t v, uint32_t n) { return n >= 32 ? 0u : v >> n; }
static inline uint32_t clz32(uint32_t v) { return v ? (uint32_t)__builtin_clz(v) : 32u; }
static inline uint32_t ffs32(uint32_t v) { return (uint32_t)__builtin_ffs((int)v); }

}  // namespace simt
#define KZG_SIMT_D static inline
