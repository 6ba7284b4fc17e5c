// Per-thread stage bodies of the decode pipeline (see kzg_kernels.cu for the pipeline overview).
// KZG_HD: the kernels call these with one GPU thread per frame / block; tests/hostsim calls the very
// same functions in a CPU loop to check the logic against libzstd without a GPU (test infrastructure).
#pragma once
#include "kzg_parse.cuh"
#include "kzg_entropy.cuh"

#if defined(__CUDA_ARCH__)
#define KZG_ATOMIC_ADD32(p, v) atomicAdd((p), (v))
#define KZG_ATOMIC_ADD64(p, v) ((u64)atomicAdd((unsigned long long*)(p), (unsigned long long)(v)))
#define KZG_ATOMIC_OR32(p, v) atomicOr((p), (v))
#else
static inline u32 kzg_host_add32(u32* p, u32 v) { const u32 o = *p; *p += v; return o; }
static inline u64 kzg_host_add64(u64* p, u64 v) { const u64 o = *p; *p += v; return o; }
static inline u32 kzg_host_or32(u32* p, u32 v) { const u32 o = *p; *p |= v; return o; }
#define KZG_ATOMIC_ADD32(p, v) kzg_host_add32((p), (v))
#define KZG_ATOMIC_ADD64(p, v) kzg_host_add64((p), (v))
#define KZG_ATOMIC_OR32(p, v) kzg_host_or32((p), (v))
#endif

// Entropy description a table owner stands for: a compressed block of the batch, the predefined
// distributions (KZG_NONE) or the formatted dictionary (KZG_DICT_OWNER).
KZG_HD const KzgFseDesc* kzg_fse_desc_of(const KzgBatch& B, u32 owner) {
    return owner == KZG_NONE ? B.fse_default : (owner == KZG_DICT_OWNER ? B.fse_dict : &B.fse[owner]);
}
KZG_HD const KzgHufDesc* kzg_huf_desc_of(const KzgBatch& B, u32 owner) {
    return owner == KZG_DICT_OWNER ? B.huf_dict : &B.huf[owner];
}

// scan: frame header + block-header walk of frame i -> KzgFrame, KzgBlock[], KzgCBlock[] stubs (stage s1)
KZG_HD void kzg_stage_scan(const KzgBatch& B, u32 i) {
    const u8* f = B.src + B.src_off[i];
    const u64 fsz = B.src_size[i];
    KzgFrame fr;
    fr.window_size = 0; fr.fcs = ~0ull; fr.out_size = 0; fr.first_block = 0; fr.n_blocks = 0;
    fr.hdr_size = 0; fr.dict_id = 0; fr.checksum_off = KZG_NONE; fr.status = KZG_OK; fr.block_max = 0; fr.xxh_lo = 0; fr.lit_garbage = 0;
    fr.fail_block = 0; fr.pad_ = 0;
    fr.end_owner[0] = fr.end_owner[1] = fr.end_owner[2] = fr.end_owner[3] = KZG_INVALID_OWNER;
    fr.end_rep[0] = 1; fr.end_rep[1] = 4; fr.end_rep[2] = 8;
    if (fsz >= 8 && (kzg_rd32(f) & 0xFFFFFFF0u) == KZG_SKIP_MAGIC) {     // skippable frame: no output
        if (8ull + kzg_rd32(f + 4) != fsz) fr.status = KZG_E_SRCSIZE_WRONG;
        B.frames[i] = fr;
        return;
    }
    KzgFrameHeader h;
    u32 err = kzg_parse_frame_header(f, fsz, h);
    if (!err && !h.single_segment && h.window_size > (1ull << 27) + 1) err = KZG_E_WINDOW_TOO_LARGE;
    if (!err && h.dict_id && h.dict_id != B.dict_id) err = KZG_E_DICT_WRONG;
    u32 nb = 0, ncb = 0;
    u64 pos = 0;
    u32 block_max = 0;
    if (!err) {
        block_max = h.window_size < KZG_BLOCK_MAX ? (u32)h.window_size : KZG_BLOCK_MAX;
        if (B.seg_mode) block_max = B.seg_block_max;
        pos = h.hdr_size;
        for (;;) {
            if (pos + 3 > fsz) { err = KZG_E_SRCSIZE_WRONG; break; }
            const u32 bh = kzg_rd24(f + pos);
            const u32 type = (bh >> 1) & 3, bsz = bh >> 3;
            if (type == 3 || bsz > block_max) { err = KZG_E_CORRUPTION; break; }
            const u32 adv = type == 1 ? 1 : bsz;
            if (pos + 3 + adv > fsz) { err = KZG_E_SRCSIZE_WRONG; break; }
            nb++; ncb += (type == 2);
            pos += 3 + adv;
            if (bh & 1) break;
        }
    }
    if (!err && h.has_checksum) {
        fr.checksum_off = (u32)pos;
        pos += 4;
        if (pos > fsz) err = KZG_E_SRCSIZE_WRONG;
    }
    if (!err && pos != fsz) err = KZG_E_SRCSIZE_WRONG;
    if (err) { fr.status = err; B.frames[i] = fr; return; }
    fr.window_size = h.window_size; fr.fcs = h.fcs; fr.hdr_size = h.hdr_size; fr.dict_id = h.dict_id;
    fr.block_max = block_max;
    const u32 first = KZG_ATOMIC_ADD32(&B.ctr->n_blocks, nb);
    const u32 cfirst = KZG_ATOMIC_ADD32(&B.ctr->n_cblocks, ncb);
    if ((u64)first + nb > B.block_cap || (u64)cfirst + ncb > B.cblock_cap) {
        KZG_ATOMIC_OR32(&B.ctr->overflow, (u64)first + nb > B.block_cap ? 1u : 2u);
        fr.status = KZG_E_INTERNAL_OVERFLOW;
        B.frames[i] = fr;
        return;
    }
    fr.first_block = first; fr.n_blocks = nb;
    B.frames[i] = fr;
    pos = h.hdr_size;
    u32 c = cfirst;
    for (u32 k = 0; k < nb; k++) {
        const u32 bh = kzg_rd24(f + pos);
        const u32 type = (bh >> 1) & 3, bsz = bh >> 3;
        KzgBlock blk;
        blk.src = B.src_off[i] + pos + 3; blk.out_off = 0; blk.frame = i; blk.csize = bsz;
        blk.regen = type == 2 ? 0 : bsz; blk.cidx = KZG_NONE; blk.type = type; blk.last = bh & 1;
        if (type == 2) {
            blk.cidx = c;
            KzgCBlock* cb = &B.cblocks[c++];
            cb->src = blk.src; cb->block = first + k; cb->frame = i; cb->csize = bsz;
        }
        B.blocks[first + k] = blk;
        pos += 3 + (type == 1 ? 1 : bsz);
    }
}

// parse: section headers + entropy descriptions of compressed block c, pool allocation
KZG_HD void kzg_stage_parse(const KzgBatch& B, u32 c) {
    KzgCBlock cb = B.cblocks[c];
    const u32 block_max = B.frames[cb.frame].block_max;
    kzg_parse_cblock(B.src + cb.src, cb.csize, block_max, cb, B.fse[c], B.huf[c]);
    cb.lit_out = 0; cb.seq_out = 0; cb.huf_owner = KZG_INVALID_OWNER;
    cb.fse_owner[0] = cb.fse_owner[1] = cb.fse_owner[2] = KZG_INVALID_OWNER;
    cb.rep_init[0] = 1; cb.rep_init[1] = 4; cb.rep_init[2] = 8;
    cb.rep_final[0] = 1; cb.rep_final[1] = 4; cb.rep_final[2] = 8;
    cb.first_in_frame = 0;
    cb.bulk_lits = 0;
    cb.lit_garbage = 0;
    if (cb.status_lit == KZG_OK && cb.lit_type >= 2) {
        const u64 sz = ((u64)cb.lit_regen + 31) & ~15ull;
        const u64 off = KZG_ATOMIC_ADD64(&B.ctr->lit_used, sz);
        if (off + sz > B.lit_cap) { KZG_ATOMIC_OR32(&B.ctr->overflow, 4u); cb.status_lit = KZG_E_INTERNAL_OVERFLOW; }
        else { cb.lit_out = off; B.hlist[KZG_ATOMIC_ADD32(&B.ctr->n_hblocks, 1u)] = c; }
    }
    if (cb.status_lit == KZG_OK && cb.status_seq == KZG_OK && cb.n_seq > 0) {
        const u64 off = KZG_ATOMIC_ADD64(&B.ctr->seq_used, (u64)cb.n_seq);
        if (off + cb.n_seq > B.seq_cap) { KZG_ATOMIC_OR32(&B.ctr->overflow, 8u); cb.status_seq = KZG_E_INTERNAL_OVERFLOW; }
        else { cb.seq_out = off; B.slist[KZG_ATOMIC_ADD32(&B.ctr->n_sblocks, 1u)] = c; }
    }
    B.cblocks[c] = cb;
}

// link: table owners of treeless-literals blocks and repeat-mode tables inside frame i
KZG_HD void kzg_stage_link(const KzgBatch& B, u32 i) {
    const KzgFrame fr = B.frames[i];
    if (fr.status != KZG_OK) return;
    // a formatted dictionary preloads the Huffman table and the three FSE tables of every frame
    const u32 start = B.dict_entropy ? KZG_DICT_OWNER : KZG_INVALID_OWNER;
    u32 huf_cur = start, fse_cur[3] = { start, start, start };
    if (B.seg_mode) { huf_cur = B.init_owner[0]; fse_cur[0] = B.init_owner[1]; fse_cur[1] = B.init_owner[2]; fse_cur[2] = B.init_owner[3]; }
    u32 seen_seq = 0;
    for (u32 k = 0; k < fr.n_blocks; k++) {
        const u32 c = B.blocks[fr.first_block + k].cidx;
        if (c == KZG_NONE) continue;
        KzgCBlock* cb = &B.cblocks[c];
        if (cb->lit_type == 2) { huf_cur = c; cb->huf_owner = c; }
        else if (cb->lit_type == 3) {
            if (huf_cur == KZG_INVALID_OWNER) { if (cb->status_lit == KZG_OK) cb->status_lit = KZG_E_DICT_CORRUPTED; }
            else cb->huf_owner = huf_cur;
        }
        if (cb->n_seq > 0 && cb->status_seq == KZG_OK) {
            for (u32 t = 0; t < 3; t++) {
                const u32 mode = (cb->modes >> (2 * t)) & 3;
                u32 owner;
                if (mode == 0) owner = KZG_NONE;
                else if (mode == 3) {
                    owner = fse_cur[t];
                    if (owner == KZG_INVALID_OWNER) cb->status_seq = KZG_E_CORRUPTION;
                } else owner = c;
                fse_cur[t] = owner;
                cb->fse_owner[t] = owner;
            }
            cb->first_in_frame = !seen_seq;
            seen_seq = 1;
        }
    }
    KzgFrame* fw = &B.frames[i];
    fw->end_owner[0] = huf_cur; fw->end_owner[1] = fse_cur[0]; fw->end_owner[2] = fse_cur[1]; fw->end_owner[3] = fse_cur[2];
}

// huffman: decode stream `st` of Huffman block h (index into hlist) with the already built table
KZG_HD void kzg_stage_huf_stream(const KzgBatch& B, KzgCBlock* cb, u32 st, const u16* l1, const u16* l2, u32 nlong, u32 maxbits) {
    const u32 regen = cb->lit_regen;
    const u8* src = B.src + cb->src + cb->lit_src;
    u8* dst = B.lit_pool + cb->lit_out;
    u32 cnt = regen, size = cb->stream_size[0];
    if (cb->n_streams == 4) {
        const u32 seg = (regen + 3) >> 2;
        for (u32 k = 0; k < st; k++) src += cb->stream_size[k];
        size = cb->stream_size[st];
        dst += st * seg;
        cnt = st < 3 ? seg : regen - 3 * seg;
    }
    const u32 e = kzg_huf_decode_stream(src, size, l1, l2, nlong, maxbits, dst, cnt);
    if (e == KZG_E_CORRUPTION || kzg_huf_verdict(cb, e)) cb->status_lit = KZG_E_CORRUPTION;
}

// fse: build the three tables of block cb into tLL/tML/tOF and run the sequence state machine
// scrLL: >= 80 B of 16-byte aligned scratch for the LL build; the ML and OF builds borrow the
// (not yet built) LL table area as scratch.
KZG_HD void kzg_stage_fse_block(const KzgBatch& B, KzgCBlock* cb, u16* tLL, u16* tML, u16* tOF, u16* scrLL,
                                const u32* llx, const u32* mlx) {
    if (cb->status_seq != KZG_OK || cb->status_lit != KZG_OK) return;
    const u32 o0 = cb->fse_owner[0], o1 = cb->fse_owner[1], o2 = cb->fse_owner[2];
    const u32 logML = kzg_fse_build_desc(tML, kzg_fse_desc_of(B, o2), 2, tLL);
    const u32 logOF = kzg_fse_build_desc(tOF, kzg_fse_desc_of(B, o1), 1, tLL);
    const u32 logLL = kzg_fse_build_desc(tLL, kzg_fse_desc_of(B, o0), 0, scrLL);
    u32 r0 = 1, r1 = 4, r2 = 8;
    if (B.dict_entropy) { r0 = B.dict_rep[0]; r1 = B.dict_rep[1]; r2 = B.dict_rep[2]; }
    if (!cb->first_in_frame) { r0 = KZG_OFF_SYM | 0; r1 = KZG_OFF_SYM | 1; r2 = KZG_OFF_SYM | 2; }
    KzgSeqOut res;
    kzg_decode_sequences(B.src + cb->src + cb->seq_src, cb->seq_size, cb->n_seq, tLL, logLL, tOF, logOF, tML, logML,
                         llx, mlx, r0, r1, r2, cb->lit_regen, B.frames[cb->frame].block_max,
                         B.seq_pool + cb->seq_out, res);
    cb->status_seq = res.status;
    cb->regen = res.regen;
    cb->lit_used = res.lit_used;
    cb->rep_final[0] = res.rep0; cb->rep_final[1] = res.rep1; cb->rep_final[2] = res.rep2;
    cb->bulk_lits = res.bulk_lits >= KZG_BULK_SUM ? 1u : 0u;
}

KZG_HD u32 kzg_rep_subst(u32 enc, const u32* rep) {
    if (!(enc & KZG_OFF_SYM)) return enc;
    const u32 base = rep[enc & 3], d = (enc & (KZG_OFF_SYM - 1)) >> 2;
    if (base >= KZG_OFF_INVALID || base <= d) return KZG_OFF_INVALID;
    return base - d;
}

// offsets: output offset of every block of frame i, repeat-offset hand-over, size checks
KZG_HD void kzg_stage_offsets(const KzgBatch& B, u32 i) {
    KzgFrame* fr = &B.frames[i];
    if (fr->status != KZG_OK) return;
    const u64 cap = B.dst_cap[i];
    u64 pos = 0;
    u32 err = KZG_OK;
    u32 rep[3] = { 1, 4, 8 };
    if (B.dict_entropy) { rep[0] = B.dict_rep[0]; rep[1] = B.dict_rep[1]; rep[2] = B.dict_rep[2]; }
    u32 k = 0;
    for (; k < fr->n_blocks; k++) {
        KzgBlock* blk = &B.blocks[fr->first_block + k];
        if (blk->type == 2) {
            KzgCBlock* cb = &B.cblocks[blk->cidx];
            err = cb->status_lit ? cb->status_lit : cb->status_seq;
            if (err) break;
            if (cb->lit_garbage && !fr->lit_garbage) { fr->lit_garbage = 1u; fr->fail_block = k; }
            u32 regen = cb->lit_regen;
            if (cb->n_seq > 0) {
                regen = cb->regen;
                cb->rep_init[0] = rep[0]; cb->rep_init[1] = rep[1]; cb->rep_init[2] = rep[2];
                const u32 a = kzg_rep_subst(cb->rep_final[0], rep), b = kzg_rep_subst(cb->rep_final[1], rep),
                          c = kzg_rep_subst(cb->rep_final[2], rep);
                rep[0] = a; rep[1] = b; rep[2] = c;
            }
            if (regen > fr->block_max) { err = KZG_E_CORRUPTION; break; }
            blk->regen = regen;
        }
        blk->out_off = pos;
        pos += blk->regen;
        if (pos > cap) { err = KZG_E_DSTSIZE_TOO_SMALL; break; }
    }
    // (libzstd compares the content size when it decodes the last block: the blocks before it have been handed over)
    if (!err && fr->fcs != ~0ull && pos != fr->fcs) { err = KZG_E_CORRUPTION; k = fr->n_blocks - 1; }
    fr->status = err;
    fr->out_size = err ? 0 : pos;
    if (err) { fr->fail_block = k; return; }
    fr->end_rep[0] = rep[0]; fr->end_rep[1] = rep[1]; fr->end_rep[2] = rep[2];
    // work list of the bulk kernel: raw / RLE blocks, literal-only blocks, long literal tails and runs
    for (u32 k = 0; k < fr->n_blocks; k++) {
        const KzgBlock* blk = &B.blocks[fr->first_block + k];
        bool bulk = blk->type != 2;
        if (!bulk) {
            const KzgCBlock* cb = &B.cblocks[blk->cidx];
            bulk = cb->n_seq == 0 || cb->lit_regen - cb->lit_used >= KZG_BULK_MIN || cb->bulk_lits;
        }
        if (bulk && blk->regen) B.blist[KZG_ATOMIC_ADD32(&B.ctr->n_bulk, 1u)] = fr->first_block + k;
    }
}
