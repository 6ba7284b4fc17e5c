/* TEST/BENCH INFRASTRUCTURE ONLY (built into oracle/_ref/).
 *
 * CPU baseline = the reference's own CPU implementation of the hot path: libzstd, the library
 * the reference's JNI wrapper calls (Wrapper.cpp:178 ZSTD_decompressStream). The Kotlin/JNI
 * layer cannot be compiled here (no JVM / jni.h / zstd.h), so this driver dlopen()s the libzstd
 * binary present in the image and calls it natively:
 *   mode 0 "parallel one-shot": one ZSTD_DCtx per thread, ZSTD_decompressDCtx per frame, static
 *          contiguous partition of the frame list (mode A);
 *   mode 1 "reference call pattern": ZSTD_decompressStream driven like
 *          SliceTransform.transform(ByteArray) (SliceTransform.kt:33-45): whole input visible,
 *          output slices of max(8192, n/10) bytes, per frame (mode B).
 * No source from /root/reference is copied.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { const void* src; size_t size; size_t pos; } zin_t;
typedef struct { void* dst; size_t size; size_t pos; } zout_t;
typedef struct {
    void* (*createDCtx)(void);
    size_t (*freeDCtx)(void*);
    size_t (*decompressDCtx)(void*, void*, size_t, const void*, size_t);
    size_t (*decompressStream)(void*, zout_t*, zin_t*);
    unsigned (*isError)(size_t);
    const char* (*versionString)(void);
} zlib_t;

typedef struct {
    zlib_t* z; int mode; const uint8_t* comp; const uint64_t* off; const uint64_t* size;
    uint64_t first, count; uint8_t* out; size_t out_stride; uint64_t bytes; int err;
    pthread_barrier_t* bar; int passes; double* t_pass; int tid;
} job_t;

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void* worker(void* a) {
    job_t* j = (job_t*)a;
    void* d = j->z->createDCtx();
    int p;
    for (p = 0; p < j->passes; p++) {
        uint64_t i;
        double t0;
        pthread_barrier_wait(j->bar);
        t0 = now();
        j->bytes = 0;
        for (i = j->first; i < j->first + j->count; i++) {
            uint8_t* dst = j->out + (i - j->first) * j->out_stride;
            if (j->mode == 0) {
                size_t n = j->z->decompressDCtx(d, dst, j->out_stride, j->comp + j->off[i], j->size[i]);
                if (j->z->isError(n)) { j->err = 1; break; }
                j->bytes += n;
            } else {
                size_t n = j->size[i], slice = n / 10 > 8192 ? n / 10 : 8192, produced = 0;
                zin_t in = { j->comp + j->off[i], n, 0 };
                for (;;) {
                    size_t cap = slice < j->out_stride - produced ? slice : j->out_stride - produced;
                    zout_t out = { dst + produced, cap, 0 };
                    size_t r = j->z->decompressStream(d, &out, &in);
                    int insufficient;
                    if (j->z->isError(r)) { j->err = 1; break; }
                    produced += out.pos;
                    insufficient = (out.pos == out.size) && r != 0;
                    if (!(in.pos != in.size || insufficient)) break;
                    if (cap == 0) { j->err = 1; break; }
                }
                j->bytes += produced;
            }
        }
        pthread_barrier_wait(j->bar);
        if (j->tid == 0) j->t_pass[p] = now() - t0;
    }
    j->z->freeDCtx(d);
    return NULL;
}

/* Decodes n_frames frames `passes` times with `threads` threads. t_pass[passes] receives the wall
 * time of each pass (first pass = warm-up by convention of the caller). Returns total
 * decompressed bytes of one pass, 0 on error. version (>= 32 bytes) receives ZSTD_versionString. */
uint64_t kzb_run(const char* libzstd_path, int mode, const uint8_t* comp, const uint64_t* off, const uint64_t* size,
                 uint64_t n_frames, size_t out_stride, int threads, int passes, double* t_pass, char* version) {
    void* h = dlopen(libzstd_path, RTLD_NOW | RTLD_LOCAL);
    zlib_t z;
    pthread_t* th;
    job_t* jobs;
    pthread_barrier_t bar;
    uint64_t total = 0, per, first = 0;
    int t, err = 0;
    if (!h) return 0;
    *(void**)&z.createDCtx = dlsym(h, "ZSTD_createDCtx");
    *(void**)&z.freeDCtx = dlsym(h, "ZSTD_freeDCtx");
    *(void**)&z.decompressDCtx = dlsym(h, "ZSTD_decompressDCtx");
    *(void**)&z.decompressStream = dlsym(h, "ZSTD_decompressStream");
    *(void**)&z.isError = dlsym(h, "ZSTD_isError");
    *(void**)&z.versionString = dlsym(h, "ZSTD_versionString");
    if (!z.createDCtx || !z.decompressDCtx || !z.decompressStream || !z.isError) return 0;
    if (version && z.versionString) { strncpy(version, z.versionString(), 31); version[31] = 0; }
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n_frames) threads = (int)(n_frames ? n_frames : 1);
    th = (pthread_t*)calloc((size_t)threads, sizeof *th);
    jobs = (job_t*)calloc((size_t)threads, sizeof *jobs);
    pthread_barrier_init(&bar, NULL, (unsigned)threads);
    per = n_frames / (uint64_t)threads;
    for (t = 0; t < threads; t++) {
        uint64_t cnt = per + ((uint64_t)t < n_frames % (uint64_t)threads ? 1 : 0);
        job_t j = { &z, mode, comp, off, size, first, cnt, (uint8_t*)malloc(cnt * out_stride + 64), out_stride, 0, 0, &bar, passes, t_pass, t };
        jobs[t] = j;
        first += cnt;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (t = 0; t < threads; t++) { pthread_join(th[t], NULL); err |= jobs[t].err; total += jobs[t].bytes; free(jobs[t].out); }
    pthread_barrier_destroy(&bar);
    free(th); free(jobs);
    return err ? 0 : total;
}
