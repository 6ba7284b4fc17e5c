/* BENCH/TEST INFRASTRUCTURE: synthetic corpora for BASELINE.json's configs (SURVEY.md section 8(d),
 * Appendix E) and a multi-threaded frame builder that compresses them with the real libzstd
 * (dlopen; the reference's compressor side stays on libzstd, SURVEY.md section 2 "zstd compressor").
 * Frame i of a corpus is generated from splitmix64(seed ^ i), so generation is parallel,
 * shardable across ranks and reproducible.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;

typedef struct { u64 s; } rng_t;
static u64 splitmix64(u64* s) {
    u64 z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static u64 rnd(rng_t* r) { return splitmix64(&r->s); }
static u32 rndn(rng_t* r, u32 n) { return (u32)((rnd(r) >> 32) * (u64)n >> 32); }   /* [0,n) */
static u32 rndr(rng_t* r, u32 lo, u32 hi) { return lo + rndn(r, hi - lo + 1); }    /* [lo,hi] */
static double rndf(rng_t* r) { return (double)(rnd(r) >> 11) * (1.0 / 9007199254740992.0); }

#define PICK(r, arr) (arr[rndn(r, (u32)(sizeof(arr) / sizeof(arr[0])))])

/* ---------------------------------------------------------------- E.1 log-like text */
static const char* LEVELS[20] = { "INFO","INFO","INFO","INFO","INFO","INFO","INFO","INFO","INFO","INFO","INFO","INFO","INFO","INFO","DEBUG","DEBUG","DEBUG","WARN","WARN","ERROR" };
static const char* COMPONENTS[10] = { "http.server","db.pool","auth.session","cache.lru","queue.worker","grpc.client","scheduler","storage.s3","metrics.push","tls.handshake" };
static const char* POOLS[3] = { "w","io","bg" };
static const char* METHODS[5] = { "GET","GET","POST","PUT","DELETE" };
static const char* RES[4] = { "items","users","orders","sessions" };
static const int STATUS[9] = { 200,200,200,201,204,304,400,404,500 };
static const char* DBPOOL[3] = { "primary","replica-1","replica-2" };
static const char* CACHEOP[5] = { "hit","hit","hit","miss","evict" };
static const int TTL[4] = { 60,300,900,3600 };
static const char* OPS[3] = { "put_object","send_batch","renew_lease" };
static const int BACKOFF[5] = { 50,100,200,400,800 };
static const char* ERRS[3] = { "timeout","connection reset by peer","503 Service Unavailable" };
static const char* QUEUES[3] = { "default","bulk","priority" };

static size_t gen_log_line(rng_t* r, u64* ts_ms, char* out) {
    int n;
    *ts_ms += rndn(r, 40);
    n = sprintf(out, "%llu.%03u %-5s [%s] (%s-%u) ", (unsigned long long)(*ts_ms / 1000), (unsigned)(*ts_ms % 1000),
                LEVELS[rndn(r, 20)], PICK(r, COMPONENTS), PICK(r, POOLS), rndn(r, 16));
    switch (rndn(r, 5)) {
        case 0:
            n += sprintf(out + n, "request completed method=%s path=/api/v%u/%s/%u status=%d bytes=%u duration_ms=%u trace=%016llx",
                         PICK(r, METHODS), rndr(r, 1, 3), PICK(r, RES), rndn(r, 100000), PICK(r, STATUS),
                         rndr(r, 64, 89999), rndr(r, 1, 899), (unsigned long long)rnd(r));
            break;
        case 1:
            n += sprintf(out + n, "connection acquired pool=%s idle=%u active=%u wait_us=%u", PICK(r, DBPOOL),
                         rndn(r, 32), rndn(r, 64), rndn(r, 5000));
            break;
        case 2:
            n += sprintf(out + n, "cache %s key=user:%u:profile ttl=%d size=%u", PICK(r, CACHEOP), rndn(r, 10000000),
                         PICK(r, TTL), rndr(r, 128, 8191));
            break;
        case 3:
            n += sprintf(out + n, "retrying operation op=%s attempt=%u backoff_ms=%d err=\"%s\"", PICK(r, OPS),
                         rndr(r, 1, 5), PICK(r, BACKOFF), PICK(r, ERRS));
            break;
        default:
            n += sprintf(out + n, "job finished id=%08x queue=%s items=%u elapsed=%u.%03us", (unsigned)rnd(r),
                         PICK(r, QUEUES), rndr(r, 1, 4999), rndn(r, 30), rndn(r, 1000));
    }
    out[n++] = '\n';
    return (size_t)n;
}

void kzc_gen_logs(u64 seed, u8* dst, size_t n) {
    rng_t r = { seed };
    u64 ts = 1760769333000ULL + (rnd(&r) % 86400000ULL);
    char line[512];
    size_t pos = 0;
    while (pos < n) {
        size_t l = gen_log_line(&r, &ts, line);
        if (l > n - pos) l = n - pos;
        memcpy(dst + pos, line, l);
        pos += l;
    }
}

/* ---------------------------------------------------------------- E.2 JSON lines */
static const char* WORDS[26] = { "alpha","bravo","charlie","delta","echo","foxtrot","golf","hotel","india","juliett","kilo","lima","mike","november","oscar","papa","quebec","romeo","sierra","tango","uniform","victor","whiskey","xray","yankee","zulu" };
static const char* EVENTS[6] = { "click","view","purchase","scroll","login","logout" };

static size_t gen_json_line(rng_t* r, char* out) {
    int n, i, ntags;
    n = sprintf(out, "{\"id\":%llu,\"ts\":\"2026-10-%02uT%02u:%02u:%02u.%03uZ\",\"user\":{\"name\":\"%s_%s\",\"age\":%u,\"verified\":%s,\"score\":%.3f},\"tags\":[",
                (unsigned long long)(rnd(r) >> 16), rndr(r, 1, 28), rndn(r, 24), rndn(r, 60), rndn(r, 60), rndn(r, 1000),
                PICK(r, WORDS), PICK(r, WORDS), rndr(r, 18, 89), rndn(r, 10) < 7 ? "true" : "false", rndf(r) * 1000.0);
    ntags = (int)rndn(r, 6);
    for (i = 0; i < ntags; i++) n += sprintf(out + n, "%s\"%s\"", i ? "," : "", PICK(r, WORDS));
    n += sprintf(out + n, "],\"geo\":{\"lat\":%.5f,\"lon\":%.5f},\"event\":\"%s\",\"value\":", rndf(r) * 180.0 - 90.0,
                 rndf(r) * 360.0 - 180.0, PICK(r, EVENTS));
    switch (rndn(r, 3)) {
        case 0: n += sprintf(out + n, "null"); break;
        case 1: n += sprintf(out + n, "%u", rndn(r, 10000)); break;
        default: n += sprintf(out + n, "%.6f", rndf(r));
    }
    n += sprintf(out + n, ",\"session\":\"%016llx%016llx\"}\n", (unsigned long long)rnd(r), (unsigned long long)rnd(r));
    return (size_t)n;
}

void kzc_gen_json(u64 seed, u8* dst, size_t n) {
    rng_t r = { seed };
    char line[1024];
    size_t pos = 0;
    while (pos < n) {
        size_t l = gen_json_line(&r, line);
        if (l > n - pos) l = n - pos;
        memcpy(dst + pos, line, l);
        pos += l;
    }
}

/* ---------------------------------------------------------------- E.3 literal-heavy mix */
static double g_cdf[256];
static pthread_once_t g_cdf_once = PTHREAD_ONCE_INIT;
static void cdf_init(void) {
    extern double pow(double, double);
    double tot = 0;
    int i;
    for (i = 0; i < 256; i++) { tot += 1.0 / pow(1.0 + i, 0.6); g_cdf[i] = tot; }   /* weights 1/(1+i)^0.6 */
    for (i = 0; i < 256; i++) g_cdf[i] /= tot;
}

void kzc_gen_literal_heavy(u64 seed, u8* dst, size_t n) {
    rng_t r = { seed };
    const double* cdf = g_cdf;
    size_t pos = 0;
    pthread_once(&g_cdf_once, cdf_init);
    while (pos < n) {
        static const size_t SEG[3] = { 16384, 32768, 65536 };
        size_t seg = SEG[rndn(&r, 3)], i;
        double kind = rndf(&r);
        if (seg > n - pos) seg = n - pos;
        if (kind < 0.55) {
            for (i = 0; i < seg; i++) {
                double u = rndf(&r);
                int lo = 0, hi = 255;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] < u) lo = mid + 1; else hi = mid; }
                dst[pos + i] = (u8)lo;
            }
        } else if (kind < 0.90) {
            for (i = 0; i + 8 <= seg; i += 8) { u64 v = rnd(&r); memcpy(dst + pos + i, &v, 8); }
            for (; i < seg; i++) dst[pos + i] = (u8)rnd(&r);
        } else {
            memset(dst + pos, (int)rndn(&r, 256), seg);
        }
        pos += seg;
    }
}

void kzc_gen_random(u64 seed, u8* dst, size_t n) {
    rng_t r = { seed };
    size_t i;
    for (i = 0; i + 8 <= n; i += 8) { u64 v = rnd(&r); memcpy(dst + i, &v, 8); }
    for (; i < n; i++) dst[i] = (u8)rnd(&r);
}

/* E.4: first half logs, second half JSON */
void kzc_gen_mixed(u64 seed, u8* dst, size_t n) {
    kzc_gen_logs(seed, dst, n / 2);
    kzc_gen_json(seed ^ 0x5DEECE66DULL, dst + n / 2, n - n / 2);
}

void kzc_gen(int kind, u64 seed, u8* dst, size_t n) {
    switch (kind) {
        case 0: kzc_gen_logs(seed, dst, n); break;
        case 1: kzc_gen_json(seed, dst, n); break;
        case 2: kzc_gen_literal_heavy(seed, dst, n); break;
        case 3: kzc_gen_random(seed, dst, n); break;
        default: kzc_gen_mixed(seed, dst, n);
    }
}

u64 kzc_frame_seed(u64 seed, u64 i) { u64 s = seed ^ i; return splitmix64(&s); }

/* ---------------------------------------------------------------- frame builder */
typedef struct { const void* src; size_t size; size_t pos; } zin_t;
typedef struct { void* dst; size_t size; size_t pos; } zout_t;
typedef struct {
    void* h;
    void* (*createCCtx)(void);
    size_t (*freeCCtx)(void*);
    size_t (*setParameter)(void*, int, int);
    size_t (*compress2)(void*, void*, size_t, const void*, size_t);
    size_t (*compressStream2)(void*, zout_t*, zin_t*, int);
    size_t (*compressBound)(size_t);
    unsigned (*isError)(size_t);
    size_t (*reset)(void*, int);
} zlib_t;

static int zlib_open(zlib_t* z, const char* path) {
    z->h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!z->h) return -1;
    *(void**)&z->createCCtx = dlsym(z->h, "ZSTD_createCCtx");
    *(void**)&z->freeCCtx = dlsym(z->h, "ZSTD_freeCCtx");
    *(void**)&z->setParameter = dlsym(z->h, "ZSTD_CCtx_setParameter");
    *(void**)&z->compress2 = dlsym(z->h, "ZSTD_compress2");
    *(void**)&z->compressStream2 = dlsym(z->h, "ZSTD_compressStream2");
    *(void**)&z->compressBound = dlsym(z->h, "ZSTD_compressBound");
    *(void**)&z->isError = dlsym(z->h, "ZSTD_isError");
    *(void**)&z->reset = dlsym(z->h, "ZSTD_CCtx_reset");
    return (z->createCCtx && z->compress2 && z->compressStream2 && z->compressBound && z->isError && z->reset) ? 0 : -2;
}

typedef struct {
    zlib_t* z;
    int kind, level, checksum, streamed, tid, nthreads;
    u64 seed, first_frame, n_frames;
    size_t frame_size, slot;
    u8* raw;      /* n_frames * frame_size (may be NULL: raw not kept) */
    u8* comp;     /* n_frames * slot */
    u64* csize;   /* n_frames */
    int err;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    void* cctx = j->z->createCCtx();
    u8* tmp = j->raw ? NULL : (u8*)malloc(j->frame_size);
    u64 i;
    for (i = (u64)j->tid; i < j->n_frames; i += (u64)j->nthreads) {
        u8* raw = j->raw ? j->raw + i * j->frame_size : tmp;
        u8* dst = j->comp + i * j->slot;
        size_t n;
        kzc_gen(j->kind, kzc_frame_seed(j->seed, j->first_frame + i), raw, j->frame_size);
        j->z->reset(cctx, 1 /* session only */);
        j->z->setParameter(cctx, 100, j->level);
        j->z->setParameter(cctx, 201, j->checksum);
        if (!j->streamed) {
            n = j->z->compress2(cctx, dst, j->slot, raw, j->frame_size);
            if (j->z->isError(n)) { j->err = 1; break; }
        } else {
            zout_t out = { dst, j->slot, 0 };
            size_t pos = 0, r;
            const size_t CH = 128 * 1024;
            while (pos < j->frame_size) {
                size_t end = pos + CH < j->frame_size ? pos + CH : j->frame_size;
                zin_t in = { raw, end, pos };
                while (in.pos < in.size) {
                    r = j->z->compressStream2(cctx, &out, &in, 0);
                    if (j->z->isError(r)) { j->err = 1; goto done; }
                }
                pos = end;
            }
            for (;;) {
                zin_t in = { raw, j->frame_size, j->frame_size };
                r = j->z->compressStream2(cctx, &out, &in, 2);
                if (j->z->isError(r)) { j->err = 1; goto done; }
                if (r == 0) break;
            }
            n = out.pos;
        }
        j->csize[i] = n;
    }
done:
    free(tmp);
    j->z->freeCCtx(cctx);
    return NULL;
}

/* Generates frames [first_frame, first_frame+n_frames) of a corpus and compresses each
 * independently. raw (n_frames*frame_size) may be NULL. comp must hold n_frames*slot bytes with
 * slot >= kzc_slot_size(frame_size). On return comp is compacted (frames back to back), csize[i]
 * holds each frame's compressed size. Returns total compressed bytes, or 0 on error. */
size_t kzc_slot_size(size_t frame_size) { return frame_size + (frame_size >> 7) + 1024; }

u64 kzc_build(const char* libzstd_path, int kind, u64 seed, u64 first_frame, u64 n_frames, size_t frame_size,
              int level, int checksum, int streamed, int nthreads, u8* raw, u8* comp, u64* csize) {
    zlib_t z;
    pthread_t th[256];
    job_t jobs[256];
    size_t slot = kzc_slot_size(frame_size);
    u64 i, total = 0;
    int t, err = 0;
    if (zlib_open(&z, libzstd_path) != 0) return 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    for (t = 0; t < nthreads; t++) {
        job_t j = { &z, kind, level, checksum, streamed, t, nthreads, seed, first_frame, n_frames, frame_size, slot, raw, comp, csize, 0 };
        jobs[t] = j;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); err |= jobs[t].err; }
    if (err) return 0;
    for (i = 0; i < n_frames; i++) {
        if (total != i * slot) memmove(comp + total, comp + i * slot, csize[i]);
        total += csize[i];
    }
    return total;
}
