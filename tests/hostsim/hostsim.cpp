// TEST INFRASTRUCTURE ONLY.  Runs the per-thread stage functions of the CUDA decoder
// (kompressor_b200/csrc/kzg_stages.cuh -- the very code the kernels execute, compiled here as host
// C++) in plain CPU loops, so scan / parse / link / Huffman / FSE / offsets logic and the packed
// sequence-record format can be checked against libzstd without a GPU.  The execute step below is
// a simple sequential stand-in for the CUDA execute kernel (which is warp-cooperative and only
// testable on a GPU).  Never linked into libkzg.so.
#define KZG_SIMT_EMU 1
#include "../../kompressor_b200/csrc/kzg_stages.cuh"
#include "../../kompressor_b200/csrc/kzg_exec2.cuh"      // the second execute design, run on the lock-step warp emulator (simt_emu.h)
#include "../../kompressor_b200/csrc/kzg_exec3.cuh"      // the third one (a thread per frame), likewise

#include <stdlib.h>
#include <string.h>
#include <vector>

static u64 xxh64(const u8* p, u64 len) {
    const u64 P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full, P3 = 0x165667B19E3779F9ull,
              P4 = 0x85EBCA77C2B2AE63ull, P5 = 0x27D4EB2F165667C5ull;
    auto rotl = [](u64 x, int r) { return (x << r) | (x >> (64 - r)); };
    auto round = [&](u64 a, u64 x) { return rotl(a + x * P2, 31) * P1; };
    const u8* end = p + len;
    u64 h;
    if (len >= 32) {
        u64 v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0 - P1;
        while (end - p >= 32) { v1 = round(v1, kzg_rd64(p)); v2 = round(v2, kzg_rd64(p + 8)); v3 = round(v3, kzg_rd64(p + 16)); v4 = round(v4, kzg_rd64(p + 24)); p += 32; }
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        h = (h ^ round(0, v1)) * P1 + P4; h = (h ^ round(0, v2)) * P1 + P4; h = (h ^ round(0, v3)) * P1 + P4; h = (h ^ round(0, v4)) * P1 + P4;
    } else h = P5;
    h += len;
    while (end - p >= 8) { h = rotl(h ^ round(0, kzg_rd64(p)), 27) * P1 + P4; p += 8; }
    if (end - p >= 4) { h = rotl(h ^ ((u64)kzg_rd32(p) * P1), 23) * P2 + P3; p += 4; }
    while (p < end) { h = rotl(h ^ ((u64)(*p) * P5), 11) * P1; p++; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

// Execute step: 0 = sequential stand-in, 1 / 2 = kzg_exec2_block on the warp emulator (lanes scheduled in forward /
// reverse order), with the ring size TILE.  What kzg_bulk_kernel does is done sequentially first, as on the GPU.
static u64 g_huf_long_max = 0, g_huf_long_sum = 0, g_huf_tables = 0;      // statistics of the two-level Huffman tables
extern "C" void kzs_huf_stats(u64* out) { out[0] = g_huf_long_max; out[1] = g_huf_long_sum; out[2] = g_huf_tables; g_huf_long_max = g_huf_long_sum = g_huf_tables = 0; }
static int g_exec_mode = 0;
static u32 g_exec_tile = 8192;
static u64 g_exec_deferred = 0, g_exec_collectives = 0;
extern "C" void kzs_set_exec(int mode, u32 tile) { g_exec_mode = mode; g_exec_tile = tile; }
extern "C" u64 kzs_exec_deferred() { return g_exec_deferred; }

struct X2Arg { const KzgBatch* B; u32 frame; u32 result; };
template <u32 TILE> static void x2_lane(void* p, u32 lane) {
    X2Arg* A = (X2Arg*)p;
    const KzgBatch& B = *A->B;
    const KzgFrame* fr = &B.frames[A->frame];
    u8* const fdst = B.dst + B.dst_off[A->frame];
    u32 r = 0;
    for (u32 k = 0; k < fr->n_blocks && r == 0; k++) {
        const KzgBlock blk = B.blocks[fr->first_block + k];
        if (blk.type != 2) continue;
        const KzgCBlock* cb = &B.cblocks[blk.cidx];
        if (cb->n_seq == 0) continue;
        if (cb->bulk_lits) { r = 2; break; }
        r = kzg_exec2_block<TILE>(B, blk, cb, fdst + blk.out_off, simt::Warp::SMEM_BASE, lane);
    }
    if (lane == 0) A->result = r;
}
static u32 x2_run_frame(const KzgBatch& B, u32 f) {
    X2Arg A{ &B, f, 0 };
    const bool rev = g_exec_mode == 2;
    switch (g_exec_tile) {
        case 4096: simt::run_warp(x2_lane<4096>, &A, KzgX2<4096>::BYTES, rev); break;
        case 16384: simt::run_warp(x2_lane<16384>, &A, KzgX2<16384>::BYTES, rev); break;
        case 32768: simt::run_warp(x2_lane<32768>, &A, KzgX2<32768>::BYTES, rev); break;
        default: simt::run_warp(x2_lane<8192>, &A, KzgX2<8192>::BYTES, rev); break;
    }
    return A.result;
}

// Execute modes 3 / 4 (5 / 6: with eager asynchronous copies): kzg_exec3_warp (lanes in forward / reverse order), ring size g_exec_tile; one emulated warp takes
// all frames of the batch by ticket.
struct X3Arg { const KzgBatch* B; };
template <u32 W> static void x3_lane(void* p, u32 lane) { kzg_exec3_warp<W>(*((X3Arg*)p)->B, simt::Warp::SMEM_BASE, lane); }
static void x3_run_batch(const KzgBatch& B) {
    X3Arg A{ &B };
    const bool rev = g_exec_mode == 4 || g_exec_mode == 6;
    simt::g_cp_eager = g_exec_mode >= 5;           // modes 5 / 6: asynchronous copies land at once (early-overwrite mistakes show)
    switch (g_exec_tile) {
        case 512: simt::run_warp(x3_lane<512>, &A, KzgX3<512>::BYTES, rev); break;
        case 2048: simt::run_warp(x3_lane<2048>, &A, KzgX3<2048>::BYTES, rev); break;
        default: simt::run_warp(x3_lane<1024>, &A, KzgX3<1024>::BYTES, rev); break;
    }
}

// What kzg_bulk_kernel writes for frame fr (kzg_bulk.cuh); with_runs: also the long literal runs of flagged blocks.
static void bulk_standin(const KzgBatch& B, const KzgFrame& fr, u8* fdst, bool with_runs) {
    for (u32 k = 0; k < fr.n_blocks; k++) {
        const KzgBlock& blk = B.blocks[fr.first_block + k];
        u8* out = fdst + blk.out_off;
        if (blk.type == 0) { memcpy(out, B.src + blk.src, blk.regen); continue; }
        if (blk.type == 1) { memset(out, B.src[blk.src], blk.regen); continue; }
        const KzgCBlock& cb = B.cblocks[blk.cidx];
        const u8* l = cb.lit_type >= 2 ? B.lit_pool + cb.lit_out : B.src + cb.src + cb.lit_src;
        if (cb.n_seq == 0) { if (cb.lit_type == 1) memset(out, l[0], blk.regen); else memcpy(out, l, blk.regen); continue; }
        const u32 trail = cb.lit_regen - cb.lit_used;
        if (trail >= KZG_BULK_MIN) { if (cb.lit_type == 1) memset(out + blk.regen - trail, l[0], trail); else memcpy(out + blk.regen - trail, l + cb.lit_used, trail); }
        if (!with_runs || !cb.bulk_lits) continue;
        u32 p_lit = 0, p_out = 0;
        for (u32 q = 0; q < cb.n_seq; q++) {
            const u64 rec = B.seq_pool[cb.seq_out + q];
            const u32 lit_end = (u32)(rec >> 28) & 0x3FFFF, ll = lit_end - p_lit;
            if (lit_end >= p_lit && ll >= KZG_BULK_MIN && p_out + ll <= blk.regen) { if (cb.lit_type == 1) memset(out + p_out, l[0], ll); else memcpy(out + p_out, l + p_lit, ll); }
            p_lit = lit_end; p_out = (u32)(rec >> 46);
        }
    }
}

extern "C" int kzs_decompress_batch(const u8* src_in, u64 src_len, const u64* src_off, const u64* src_size, u64 n,
                                    u8* dst, const u64* dst_off, const u64* dst_cap, u32* status, u64* out_size,
                                    u64* seq_dump, u64 seq_dump_cap, u8* lit_dump, u64 lit_dump_cap) {
    // padded, 16-byte aligned copy of the source (the device contract: aligned chunks around data are readable)
    std::vector<u8> srcbuf(src_len + 64 + 16);
    u8* src = srcbuf.data() + 16;
    src = (u8*)(((uintptr_t)src + 15) & ~(uintptr_t)15);
    memcpy(src, src_in, src_len);
    u64 total_dst = 0;
    for (u64 i = 0; i < n; i++) total_dst += dst_cap[i];
    const u64 block_cap = n + src_len / 3 + 16;
    std::vector<KzgFrame> frames(n);
    std::vector<KzgBlock> blocks(block_cap);
    std::vector<KzgCBlock> cblocks(block_cap);
    std::vector<KzgFseDesc> fse(block_cap);
    std::vector<KzgHufDesc> huf(block_cap);
    std::vector<u32> hlist(block_cap), slist(block_cap), blist(block_cap), dlist(n + 1);
    std::vector<u8> lit(total_dst + 64 * block_cap + 4096 + 16);
    std::vector<u64> seq(total_dst / 3 + 4096);
    KzgCounters ctr;
    memset(&ctr, 0, sizeof ctr);
    KzgFseDesc def;
    kzg_default_fse_desc(def);
    u32 codes[96];
    kzg_code_tables(codes, codes + 36);
    KzgBatch B;
    memset(&B, 0, sizeof B);
    B.src = src; B.src_off = src_off; B.src_size = src_size; B.dst = dst; B.dst_off = dst_off; B.dst_cap = dst_cap;
    B.status = status; B.out_size = out_size; B.n_frames = (u32)n; B.tile_bytes = KZG_BLOCK_MAX + 64;
    B.frames = frames.data(); B.blocks = blocks.data(); B.cblocks = cblocks.data(); B.fse = fse.data(); B.huf = huf.data();
    B.hlist = hlist.data(); B.slist = slist.data(); B.blist = blist.data(); B.dlist = dlist.data();
    B.lit_pool = (u8*)(((uintptr_t)lit.data() + 15) & ~(uintptr_t)15); B.seq_pool = seq.data(); B.ctr = &ctr;
    B.fse_default = &def; B.code_tables = codes; B.lit_cap = lit.size() - 32; B.seq_cap = seq.size();
    B.block_cap = (u32)block_cap; B.cblock_cap = (u32)block_cap;
    for (u32 i = 0; i < n; i++) kzg_stage_scan(B, i);
    if (ctr.overflow) return -1;
    for (u32 c = 0; c < ctr.n_cblocks; c++) kzg_stage_parse(B, c);
    for (u32 i = 0; i < n; i++) kzg_stage_link(B, i);
    // huffman: 4 "lanes" per block build the two-level table, then each stream decodes
    alignas(16) static u16 l1[1u << KZG_HUF_L1_BITS];
    alignas(16) static u16 l2[1u << KZG_HUF_MAXBITS];
    for (u32 h = 0; h < ctr.n_hblocks; h++) {
        KzgCBlock* cb = &cblocks[hlist[h]];
        if (cb->status_lit != KZG_OK || cb->huf_owner == KZG_INVALID_OWNER) continue;
        const KzgHufDesc* hd = kzg_huf_desc_of(B, cb->huf_owner);
        alignas(16) KzgHufScratch sc;
        for (u32 k = 0; k < (1u << KZG_HUF_L1_BITS); k++) l1[k] = 0xDEAD;
        for (u32 k = 0; k < (1u << KZG_HUF_MAXBITS); k++) l2[k] = 0xDEAD;
        for (u32 st = 0; st < 4; st++) kzg_huf_build_phase1(&sc, hd->weight, hd->nsym, st);
        const u32 nlong = kzg_huf_long_entries(&sc, hd->maxbits);
        if (nlong > g_huf_long_max) g_huf_long_max = nlong;
        g_huf_long_sum += nlong; g_huf_tables++;
        for (u32 st = 0; st < 4; st++) kzg_huf_build_phase2(&sc, hd->maxbits, st);
        for (u32 st = 0; st < 4; st++) kzg_huf_build_phase3(l1, l2, &sc, hd->nsym, hd->maxbits, st);
        for (u32 st = 0; st < cb->n_streams; st++) kzg_stage_huf_stream(B, cb, st, l1, l2, nlong, hd->maxbits);
    }
    alignas(16) static u16 t3[1280];
    alignas(16) static u16 scrLL[40];
    for (u32 s = 0; s < ctr.n_sblocks; s++)
        kzg_stage_fse_block(B, &cblocks[slist[s]], t3, t3 + 512, t3 + 1024, scrLL, codes, codes + 36);
    for (u32 i = 0; i < n; i++) kzg_stage_offsets(B, i);
    if (ctr.overflow) return -2;
    // optional dumps (first compressed block order) for stage-by-stage comparison
    if (seq_dump) { u64 m = ctr.seq_used < seq_dump_cap ? ctr.seq_used : seq_dump_cap; memcpy(seq_dump, seq.data(), m * 8); }
    if (lit_dump) { u64 m = ctr.lit_used < lit_dump_cap ? ctr.lit_used : lit_dump_cap; memcpy(lit_dump, B.lit_pool, m); }
    const bool x3 = g_exec_mode >= 3 && g_exec_mode <= 6;
    if (x3) {
        for (u32 f = 0; f < n; f++) if (!frames[f].status) bulk_standin(B, frames[f], dst + dst_off[f], true);
        ctr.ticket_exec = 0; ctr.n_defer = 0;
        x3_run_batch(B);
        g_exec_deferred += ctr.n_defer;
    }
    // sequential execute + finish
    for (u32 f = 0; f < n; f++) {
        KzgFrame& fr = frames[f];
        u32 st = fr.status;
        u8* fdst = dst + dst_off[f];
        bool emulated = x3;
        if (g_exec_mode != 0 && !x3 && !st) {
            // what kzg_bulk_kernel writes (kzg_bulk.cuh), then the warp-emulated execute kernel
            for (u32 k = 0; k < fr.n_blocks; k++) {
                const KzgBlock& blk = blocks[fr.first_block + k];
                u8* out = fdst + blk.out_off;
                if (blk.type == 0) { memcpy(out, src + blk.src, blk.regen); continue; }
                if (blk.type == 1) { memset(out, src[blk.src], blk.regen); continue; }
                const KzgCBlock& cb = cblocks[blk.cidx];
                const u8* l = cb.lit_type >= 2 ? B.lit_pool + cb.lit_out : src + cb.src + cb.lit_src;
                if (cb.n_seq == 0) { if (cb.lit_type == 1) memset(out, l[0], blk.regen); else memcpy(out, l, blk.regen); continue; }
                const u32 trail = cb.lit_regen - cb.lit_used;
                if (trail >= KZG_BULK_MIN) { if (cb.lit_type == 1) memset(out + blk.regen - trail, l[0], trail); else memcpy(out + blk.regen - trail, l + cb.lit_used, trail); }
            }
            const u32 r = x2_run_frame(B, f);
            if (r == 1) st = KZG_E_CORRUPTION;
            else if (r == 2) g_exec_deferred++;
            else emulated = true;
        }
        for (u32 k = 0; !st && !emulated && k < fr.n_blocks; k++) {
            const KzgBlock& blk = blocks[fr.first_block + k];
            u8* out = fdst + blk.out_off;
            if (blk.type == 0) memcpy(out, src + blk.src, blk.regen);
            else if (blk.type == 1) memset(out, src[blk.src], blk.regen);
            else {
                const KzgCBlock& cb = cblocks[blk.cidx];
                const u8* l = cb.lit_type >= 2 ? B.lit_pool + cb.lit_out : src + cb.src + cb.lit_src;
                u32 lp = 0, op = 0;
                for (u32 q = 0; q < cb.n_seq && !st; q++) {
                    const u64 rec = B.seq_pool[cb.seq_out + q];
                    const u32 lit_end = (u32)(rec >> 28) & 0x3FFFF, out_end = (u32)(rec >> 46);
                    u32 off = kzg_rep_subst((u32)rec & KZG_OFF_MASK, cb.rep_init);
                    const u32 ll = lit_end - lp;
                    for (u32 j = 0; j < ll; j++) out[op + j] = cb.lit_type == 1 ? l[0] : l[lp + j];
                    op += ll; lp = lit_end;
                    const u32 ml = out_end - op;
                    if (off == 0 || off >= KZG_OFF_INVALID || (u64)off > blk.out_off + op) { st = KZG_E_CORRUPTION; break; }
                    for (u32 j = 0; j < ml; j++) out[op + j] = out[(i64)op + j - (i64)off];
                    op += ml;
                }
                if (!st) for (u32 j = lp; j < cb.lit_regen; j++) out[op++] = cb.lit_type == 1 ? l[0] : l[j];
                if (!st && op != blk.regen) st = KZG_E_GENERIC;
            }
        }
        if (!st && fr.lit_garbage) st = fr.checksum_off != KZG_NONE ? KZG_E_CHECKSUM_WRONG : KZG_E_CORRUPTION;
        if (!st && fr.checksum_off != KZG_NONE) {
            if ((u32)xxh64(fdst, fr.out_size) != kzg_rd32(src + src_off[f] + fr.checksum_off)) st = KZG_E_CHECKSUM_WRONG;
        }
        status[f] = st;
        out_size[f] = st ? 0 : fr.out_size;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Compress side: a sequential stand-in for the warp-cooperative match kernel (same record format, same literal
// placement), then the very encode stage the GPU runs (kzg_enc_stage_encode), then the finish step in plain C.
#include "../../kompressor_b200/csrc/kzg_enc.cuh"

extern "C" int kzs_compress_batch(const u8* src, const u64* src_off, const u64* src_size, u64 n, u8* dst, const u64* dst_off,
                                  const u64* dst_cap, int flags, u32* status, u64* out_size) {
    static KzgEncTables tables;
    kzg_enc_build_tables(tables);
    u64 max_chunk = 0;
    for (u64 i = 0; i < n; i++) if (src_size[i] > max_chunk) max_chunk = src_size[i];
    if (max_chunk > KZG_BLOCK_MAX) max_chunk = KZG_BLOCK_MAX;
    const u32 seq_cap = KZG_ENC_WARPS * (u32)(max_chunk / 6 + 64);     // everything goes to the first part's region
    std::vector<u64> seq((size_t)n * seq_cap);
    std::vector<KzgEncChunk> chunks(n);
    KzgEncBatch B;
    memset(&B, 0, sizeof B);
    B.src = src; B.src_off = src_off; B.src_size = src_size; B.dst = dst; B.dst_off = dst_off; B.dst_cap = dst_cap;
    B.status = status; B.out_size = out_size; B.n = (u32)n; B.flags = (u32)flags;
    B.seq = seq.data(); B.seq_cap = seq_cap; B.chunks = chunks.data(); B.tab = &tables;
    std::vector<u32> table(1u << 12);
    for (u32 i = 0; i < n; i++) {
        KzgEncChunk& ck = chunks[i];
        for (u32 q = 0; q < KZG_ENC_WARPS; q++) ck.n_seq[q] = 0;
        ck.lit_size = 0; ck.state = 0; ck.total = 0; ck.lit_section = 0;
        const u32 len = (u32)src_size[i];
        if (src_size[i] > KZG_BLOCK_MAX || dst_cap[i] < kzg_enc_bound(len) || len < 64) continue;
        const u8* s = src + src_off[i];
        u8* lit = dst + dst_off[i] + kzg_enc_header_size(len) + 6;
        u64* rec = seq.data() + (size_t)i * seq_cap;
        std::fill(table.begin(), table.end(), 0xFFFFFFFFu);
        const u32 limit = len - 12;
        u32 pos = 0, anchor = 0;
        while (pos < limit && ck.n_seq[0] < seq_cap / KZG_ENC_WARPS) {
            const u32 v = kzg_rd32(s + pos), h = (v * 2654435761u) >> (32 - 12);
            const u32 cand = table[h];
            table[h] = pos;
            if (cand == 0xFFFFFFFFu || kzg_rd32(s + cand) != v) { pos++; continue; }
            u32 ml = 4;
            while (pos + ml < len && s[cand + ml] == s[pos + ml]) ml++;
            const u32 ll = pos - anchor;
            memcpy(lit + ck.lit_size, s + anchor, ll);
            rec[ck.n_seq[0]++] = (u64)ll | ((u64)ml << 18) | ((u64)(pos - cand) << 36);
            ck.lit_size += ll;
            anchor = pos = pos + ml;
        }
        memcpy(lit + ck.lit_size, s + anchor, len - anchor);
        ck.lit_size += len - anchor;
        kzg_enc_write_lit_header(lit - 3, 0, ck.lit_size);               // raw literals section, in place
        ck.lit_section = 3 + ck.lit_size;
    }
    for (u32 i = 0; i < n; i++) kzg_enc_stage_encode(B, i);
    for (u32 i = 0; i < n; i++) {
        const KzgEncChunk& ck = chunks[i];
        if (ck.state == KZG_ENC_ST_ERROR) continue;
        const u32 len = (u32)src_size[i];
        const u8* s = src + src_off[i];
        u8* d = dst + dst_off[i];
        u32 total = ck.total;
        if (ck.state != KZG_ENC_ST_COMPRESSED) {
            const u32 hdr = kzg_enc_header_size(len);
            kzg_wr24(d + hdr, 1u | (len << 3));
            memcpy(d + hdr + 3, s, len);
            total = hdr + 3 + len;
        }
        if (flags & KZG_ENC_FLAG_CHECKSUM) { kzg_wr32(d + total, (u32)xxh64(s, len)); total += 4; }
        status[i] = KZG_OK;
        out_size[i] = total;
    }
    return 0;
}

// FSE-coded Huffman weights: the encoder's description of `n` weights (the last symbol's implicit one is recomputed by
// the reader), read back by the decoder's own parser.  Returns the description size (0: encoder declined), -1 on mismatch.
extern "C" int kzs_weights_roundtrip(const u8* w, u32 n) {
    u8 buf[300];
    KzgEncWeightScratch sc;
    const u32 sz = kzg_enc_compress_weights(w, n, buf + 1, 127, sc);
    if (sz == 0 || sz >= 128) return 0;
    buf[0] = (u8)sz;
    u8 back[260];
    u32 nsym = 0, maxbits = 0;
    const u32 used = kzg_read_huf_weights(buf, 1 + sz, back, nsym, maxbits);
    if (used != 1 + sz || nsym != n + 1) return -1;
    for (u32 k = 0; k < n; k++) if (back[k] != w[k]) return -1;
    return (int)sz;
}
