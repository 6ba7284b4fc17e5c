// Throughput of the streaming shim (layer B of include/kzg.h) driven from native threads, in the reference's
// call patterns (no Python in the timed loops):
//   transform  SliceTransform.transform(ByteArray), SliceTransform.kt:33-45: whole input in the first call,
//              output slices of max(8192, n/10)
//   pipe xT    SliceTransformRawSource.readAtMostTo, SliceTransformRawSource.kt:32-55: 8 KiB in / 8 KiB out,
//              T threads with one handle each (one per HTTP body)
// Build + run: scripts/gpu_stream_bench.py --native  (writes the corpus file, compiles this file, runs it)
//   g++ -O2 -std=c++17 -pthread scripts/stream_bench.cpp -Iinclude -Lkompressor_b200 -lkzg -o gpurun_out/stream_bench
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "kzg.h"

using clk = std::chrono::steady_clock;
static double ms_since(clk::time_point t) { return std::chrono::duration<double, std::milli>(clk::now() - t).count(); }

struct Result { size_t out = 0; double first_ms = 0; bool ok = true; };

// Drives one handle the way the Kotlin loops do. in_slice = 0: everything in the first call.
static Result run_stream(const uint8_t* comp, size_t n, size_t in_slice, size_t out_slice, uint8_t* out, size_t out_cap) {
    Result r;
    kzg_DCtx* d = kzg_createDCtx();
    if (!d) { r.ok = false; return r; }
    size_t ip = 0, op = 0;
    bool first = true;
    while (ip < n) {
        const size_t take = in_slice == 0 ? n - ip : (in_slice < n - ip ? in_slice : n - ip);
        kzg_inBuffer ib = { comp + ip, take, 0 };
        bool insufficient = false;
        do {                                                 // SliceTransform.kt:39-45: while (input.remainingRead != 0 || output.insufficient)
            const size_t room = out_slice < out_cap - op ? out_slice : out_cap - op;
            kzg_outBuffer ob = { out + op, room, 0 };
            const auto t0 = clk::now();
            const size_t res = kzg_decompressStream(d, &ob, &ib);
            if (first) { r.first_ms = ms_since(t0); first = false; }
            if (kzg_isError(res)) { fprintf(stderr, "error: %s\n", kzg_getErrorName(res)); r.ok = false; kzg_freeDCtx(d); return r; }
            op += ob.pos;
            insufficient = ob.pos == room && res != 0;       // output.insufficient = output.isFull && result != 0
        } while (ib.pos < ib.size || insufficient);
        ip += take;
    }
    kzg_freeDCtx(d);
    r.out = op;
    return r;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: stream_bench corpus.bin frames_per_pipe_handle\n"); return 2; }
    // corpus.bin: u64 n_frames, u64 comp_bytes, u64 raw_bytes, u64 off[n+1], comp bytes, raw bytes
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("corpus"); return 2; }
    uint64_t hdr[3];
    if (fread(hdr, 8, 3, f) != 3) return 2;
    const uint64_t nf = hdr[0], cb = hdr[1], rb = hdr[2];
    std::vector<uint64_t> off(nf + 1);
    std::vector<uint8_t> comp(cb), raw(rb);
    if (fread(off.data(), 8, nf + 1, f) != nf + 1 || fread(comp.data(), 1, cb, f) != cb || fread(raw.data(), 1, rb, f) != rb) return 2;
    fclose(f);
    const size_t per = strtoull(argv[2], nullptr, 10);
    std::vector<uint8_t> out(rb + 64);
    run_stream(comp.data(), off[8 < nf ? 8 : nf], 0, 65536, out.data(), out.size());          // warm-up
    for (int rep = 0; rep < 4; rep++) {
        const auto t0 = clk::now();
        const Result r = run_stream(comp.data(), cb, 0, cb / 10 > 8192 ? cb / 10 : 8192, out.data(), out.size());
        const double ms = ms_since(t0);
        const bool same = r.ok && r.out == rb && memcmp(out.data(), raw.data(), rb) == 0;
        printf("transform %llu frames: %.2f ms (first call %.2f ms)  %.1f MB/s  parity %s\n", (unsigned long long)nf, ms, r.first_ms,
               rb / ms / 1e3, same ? "ok" : "MISMATCH");
        if (!same) return 1;
    }
    {   // layer A on the same frames: where the first call's time goes (pageable vs pinned source, pinned destination)
        kzg_ctx* c = nullptr;
        if (kzg_ctx_create(&c, -1) != 0) return 1;
        std::vector<uint64_t> o(off.begin(), off.begin() + nf), sz(nf), doff(nf), cap(nf), osz(nf);
        std::vector<uint32_t> st(nf);
        const uint64_t fr = rb / nf;
        for (uint64_t i = 0; i < nf; i++) { sz[i] = off[i + 1] - off[i]; doff[i] = i * fr; cap[i] = fr; }
        uint8_t* psrc = (uint8_t*)kzg_host_alloc(cb + 64);
        uint8_t* pdst = (uint8_t*)kzg_host_alloc(rb + 64);
        memcpy(psrc, comp.data(), cb);
        for (int mode = 0; mode < 3; mode++) {
            const uint8_t* s = mode == 0 ? comp.data() : psrc;
            uint8_t* dd = mode == 2 ? out.data() : pdst;
            double best = 1e30;
            for (int rep = 0; rep < 4; rep++) {
                const auto t0 = clk::now();
                if (kzg_decompress_batch_host(c, s, o.data(), sz.data(), dd, doff.data(), cap.data(), nf, st.data(), osz.data()) != 0) return 1;
                const double ms = ms_since(t0);
                if (ms < best) best = ms;
            }
            printf("layer A, %s source -> %s destination: %.2f ms\n", mode == 0 ? "pageable" : "pinned", mode == 2 ? "pageable" : "pinned", best);
        }
        const auto t0 = clk::now();
        memcpy(out.data(), pdst, rb);
        printf("memcpy of the output to a pageable buffer: %.2f ms\n", ms_since(t0));
        kzg_host_free(psrc); kzg_host_free(pdst);
        kzg_ctx_destroy(c);
    }
    const size_t small_c = off[per < nf ? per : nf], small_r = (size_t)(rb / nf) * (per < nf ? per : nf);
    for (int T : { 1, 4, 16, 64, 256 }) {
        std::vector<std::vector<uint8_t>> outs(T, std::vector<uint8_t>(small_r + 64));
        std::vector<Result> res(T);
        double best = 1e30;
        for (int rep = 0; rep < 3; rep++) {
            std::vector<std::thread> th;
            const auto t0 = clk::now();
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t] { res[t] = run_stream(comp.data(), small_c, 8192, 8192, outs[t].data(), outs[t].size()); });
            for (auto& x : th) x.join();
            const double ms = ms_since(t0);
            if (ms < best) best = ms;
        }
        bool same = true;
        for (int t = 0; t < T; t++) same = same && res[t].ok && res[t].out == small_r && memcmp(outs[t].data(), raw.data(), small_r) == 0;
        printf("pipe 8K/8K, %3d handles x %zu frames: %.2f ms  %.1f MB/s  parity %s\n", T, per, best, (double)small_r * T / best / 1e3,
               same ? "ok" : "MISMATCH");
        if (!same) return 1;
    }
    return 0;
}
