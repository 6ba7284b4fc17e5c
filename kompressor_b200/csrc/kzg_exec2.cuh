// Execute stage (s5, ZSTD_execSequence) for sm_100a, second design: one warp per frame, OUTPUT-centric.
//
// The first design (kzg_exec.cuh) gave every lane one sequence; its copy loops ran for the longest sequence of the
// batch (21 of 32 lanes busy), ragged ends were byte stores, and a batch of log text needed 1.9 dependency rounds.
// Here a batch of 32 sequences is only *described* by its lanes; the bytes are produced by passes over the output,
// one aligned 32-bit word per lane, 128 contiguous bytes per pass:
//   * the owner of a word is the sequence that contains the word's last byte.  Owners are found through a byte map
//     (one mark per sequence at its first owned word) + ballot + count-leading-zeros, no search;
//   * a word has at most three sources -- the last 1..3 bytes of the previous sequence's match (a match is >= 3 bytes,
//     so never more than one earlier sequence), the owner's literals, the owner's match -- merged in registers with
//     byte masks: no byte stores, no ragged ends, every lane busy whatever the sequence lengths are (long matches and
//     long literal runs are just more passes);
//   * dependencies are resolved per pass: a word goes when its match sources lie below the first pending word of the
//     pass (on log text 1.05 rounds per pass); the first pending word can always go, byte by byte; short-period
//     overlapping matches (runs) read the period instead of their own output;
//   * the output window is a RING in shared memory indexed by the low bits of the global address (no slide; its first
//     16 bytes are mirrored behind the end so unaligned word reads never wrap), flushed by ONE bulk asynchronous copy
//     shared -> global per ~0.5 KiB (cp.async.bulk, SASS UBLKCP) instead of per-lane vector stores;
//   * match sources that have left the ring are staged from global memory by cp.async one batch AHEAD (the next batch
//     is unpacked and its copies issued before the current one is executed); literals go through a small ring filled
//     the same way.
// Frames that need what this kernel leaves out (dictionary history, blocks whose long literal runs were written by the
// bulk kernel) are put on the deferred list and executed by the first design's full variant.
// What libzstd does here is ZSTD_execSequence, reached from the reference through ZSTD_decompressStream
// (kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:178).
#pragma once
#include "kzg_stages.cuh"
#include "kzg_simt.cuh"

#define KZG_X2_LR 1024u             // literal ring (bytes)
#define KZG_X2_MST 80u              // per-lane match staging: the 16-byte chunks covering <= 64 source bytes
#define KZG_X2_OWN 256u             // words per owner-map chunk
#define KZG_X2_STG 0x80000000u      // source descriptor: staged (address = warp base + position + delta) instead of ring (offset)
#define KZG_X2_PER 0x40000000u      // source descriptor (ring kind): short-period overlapping match
#define KZG_X2_OFF 0x07FFFFFFu      // source descriptor (ring kind): the offset
#define KZG_X2_BIAS 0x100000u
#define KZG_X2_LLHUGE 0x3FFFu       // literal length field of a descriptor: this many or more -> the full value is in FLD2
#define KZG_X2_FLUSH_MIN 512
#define KZG_X2_WARPS 2

template <u32 TILE> struct KzgX2 {
    static constexpr u32 RING = 0;                                  // TILE + 16 (mirror of the first 16 bytes)
    static constexpr u32 MST = TILE + 16;                           // 2 buffers x 32 lanes x 80
    static constexpr u32 LIT = MST + 2 * 32 * KZG_X2_MST;           // KZG_X2_LR + 16 (mirror)
    static constexpr u32 FLD = LIT + KZG_X2_LR + 16;                // 32 x 16: {a | ll << 18, match source, previous match source, literal delta}
    static constexpr u32 FLD2 = FLD + 512;                          // 32 x 4: mdst of sequences with huge literal runs
    static constexpr u32 OWN = FLD2 + 128;                          // owner map of the current chunk
    static constexpr u32 SCR = OWN + KZG_X2_OWN;                    // 32: scratch words (block start, carried match tail)
    static constexpr u32 BYTES = SCR + 32;
};

namespace kzg_x2 {

KZG_SIMT_D i32 mn_i(i32 a, i32 b) { return a < b ? a : b; }
KZG_SIMT_D i32 mx_i(i32 a, i32 b) { return a > b ? a : b; }
KZG_SIMT_D u32 mn_u(u32 a, u32 b) { return a < b ? a : b; }
KZG_SIMT_D u32 mx_u(u32 a, u32 b) { return a > b ? a : b; }

// 32 bits at any shared address: two aligned words + funnel shift (one word past may be read; callers keep it mapped).
KZG_SIMT_D u32 ldu32(u32 sa) {
    const u32 A = sa & ~3u;
    return simt::funnel_r(simt::lds32(A), simt::lds32(A + 4), (sa & 3u) << 3);
}
// the same from global memory (history written earlier, complete)
KZG_SIMT_D u32 ldu32_hist(const u8* base16, i32 pos) {
    const u32* p = reinterpret_cast<const u32*>(base16 + (pos & ~3));
    return simt::funnel_r(simt::ldg32_hist(p), simt::ldg32_hist(p + 1), ((u32)pos & 3u) << 3);
}
KZG_SIMT_D u32 lowmask(u32 nbytes) { return simt::shl_clamp(1u, nbytes << 3) - 1u; }     // nbytes in [0, 4]

struct Batch {          // what prepare() hands to execute(): per-lane registers + warp-uniform values
    u32 a, ll, bend, desc, litpos;
    bool glob;
    u32 nvalid, be;
    bool litdirect, has_glob, has_huge;
};

}  // namespace kzg_x2

// Executes one compressed block with sequences. Returns (uniformly) 0 = done, 1 = the block is corrupt, 2 = the frame
// needs the first design's full variant.  sb: shared address of this warp's KzgX2<TILE>::BYTES.
template <u32 TILE>
KZG_SIMT_D u32 kzg_exec2_block(const KzgBatch& B, const KzgBlock& blk, const KzgCBlock* cb, u8* out, u32 sb, u32 lane) {
    using namespace simt;
    using namespace kzg_x2;
    typedef KzgX2<TILE> L;
    static_assert(TILE >= 4096 && (TILE & (TILE - 1)) == 0, "ring size");
    constexpr int KREAD = TILE >= 8192 ? 1 : 0;        // bulk groups besides the new one whose ring reads may be outstanding
    const u32 nseq = cb->n_seq, lit_regen = cb->lit_regen, lit_type = cb->lit_type;
    const u8* lit = lit_type >= 2 ? B.lit_pool + cb->lit_out : B.src + cb->src + cb->lit_src;
    const bool lfill = lit_type == 1;
    const u32 rle = lfill ? ldg8_ro(lit) * 0x01010101u : 0u;
    const u32 lalign = (u32)((uintptr_t)lit & 15);
    const u8* lit16 = lit - lalign;
    const u64* seqs = B.seq_pool + cb->seq_out;
    const u32 ri0 = cb->rep_init[0], ri1 = cb->rep_init[1], ri2 = cb->rep_init[2];
    const u64 hist64 = blk.out_off + B.dict_size;
    const u32 hist_before_block = hist64 > 0x40000000ull ? 0x40000000u : (u32)hist64;       // clamped for 32-bit compares
    const u32 frame_before = blk.out_off > 0x40000000ull ? 0x40000000u : (u32)blk.out_off;
    // Positions are "shifted": X = block position + sh, so that X and the global address agree modulo 16 and X = 0 is
    // the 16-byte aligned address `gal`.  Ring slot of X: (rb + X) mod TILE = the global address modulo TILE.
    const u32 sh = (u32)((uintptr_t)out & 15);
    u8* const gal = out - sh;
    const u32 rb = (u32)((uintptr_t)gal) & (TILE - 1u);
    const u32 ring = sb + L::RING, litring = sb + L::LIT, fld = sb + L::FLD, fld2 = sb + L::FLD2, own = sb + L::OWN;
    const u32 scr_tail = sb + L::SCR + 16;
    // the virtual last sequence: the literals after the last real sequence (unless the bulk kernel wrote them) plus
    // padding up to the next word boundary, no match -- so that every word of the block has an owner
    const u32 trail = lit_regen - cb->lit_used;
    const u32 out_last = cb->regen - trail;
    const u32 trail_v = trail < KZG_BULK_MIN ? trail : 0u;
    const i32 cur_end = (i32)(sh + out_last + trail_v);
    const u32 pad = (0u - (u32)cur_end) & 3u;
    const u64 vrec = kzg_pack_seq(1u, cb->lit_used + trail_v + pad, out_last + trail_v + pad);
    const u32 ntot = nseq + 1u;

    i32 F = (i32)sh;            // positions below F have been handed to a flush
    i32 Fdone = (i32)sh;        // positions below Fdone are complete in global memory
    i32 S1 = (i32)sh, S2 = (i32)sh, S3 = (i32)sh, S4 = (i32)sh;     // where the four newest bulk groups start (they may still be in flight)
    const i32 Vlo = (i32)sh;    // the ring holds nothing below Vlo (earlier blocks are read from global memory)
    u32 carry_lo = 0, carry_hi = 0;                     // last record of the previous batch
    u32 carryP = KZG_X2_STG | ((L::SCR - (sh & ~3u) + KZG_X2_BIAS) & 0xFFFFFFu);   // "previous match" of the first sequence: scratch
    bool carry_glob = false;
    u32 Lf = 0;                 // literal ring: aligned stream positions below Lf have been requested
    u32 l0_cur = 0;             // first literal of the batch that executes next
    u32 status = 0;
    u64 rec_pref = lane < nseq ? ldg64_stream(seqs + lane) : 0ull;

    auto raddr = [&](i32 X) -> u32 { return ring + ((rb + (u32)X) & (TILE - 1u)); };

    // ---- flush [F, upto) (all of it final in the ring): head bytes up to a 16-byte boundary, one bulk copy, and with
    // `all` the ragged tail.  Between two flushes the output advances by at most TILE / 8 + one owner-map chunk (1 KiB);
    // the newest KREAD + 1 groups (one being issued) may still have to read the ring, which is safe while
    // (KREAD + 2) * (TILE / 8 + 1 KiB) < TILE.
    auto flush = [&](i32 upto, bool all) {
        if (upto <= F) return;
        if (F & 15) {
            const i32 nh = 16 - (F & 15);
            if (upto - F >= nh) {
                if ((i32)lane < nh) stg8(gal + F + (i32)lane, lds8(raddr(F + (i32)lane)));
                F += nh;
                fence_block();
            }
        }
        if (!(F & 15)) {
            const i32 n16 = (upto - F) & ~15;
            if (n16 > 0) {
                fence_async();
                w_sync();
                if (lane == 0) {
                    // A bulk store takes microseconds to complete, so several stay in flight: up to four older groups may
                    // still be writing, and of those only the newest two may still have to read the ring.
                    bulk_wait<4>();
                    bulk_wait_read<KREAD>();
                    const u32 slot = (rb + (u32)F) & (TILE - 1u);
                    const u32 n1 = mn_u((u32)n16, TILE - slot);
                    bulk_s2g(gal + F, ring + slot, n1);
                    if ((u32)n16 > n1) bulk_s2g(gal + F + (i32)n1, ring, (u32)n16 - n1);
                    bulk_commit();
                }
                w_sync();
                Fdone = S4;         // everything before the fourth newest group is complete
                S4 = S3; S3 = S2; S2 = S1; S1 = F;
                F += n16;
            }
        }
        if (all && F < upto) {
            const i32 nt = upto - F;
            if ((i32)lane < nt) stg8(gal + F + (i32)lane, lds8(raddr(F + (i32)lane)));
            F = upto;
            fence_block();
        }
    };

    // ---- prepare: unpack the records of batch b0, check them, classify and stage match sources, fill the literal ring
    auto prepare = [&](u32 b0, u32 buf, Batch& P) {
        const u32 i = b0 + lane;
        const u32 nvalid = mn_u(32u, ntot - b0), last = nvalid - 1u;
        const u64 rec = i < nseq ? rec_pref : vrec;
        rec_pref = (i + 32u < nseq) ? ldg64_stream(seqs + i + 32u) : 0ull;
        u32 rlo = (u32)rec, rhi = (u32)(rec >> 32);
        const u32 llo = w_shfl(rlo, last), lhi = w_shfl(rhi, last);
        if (lane >= nvalid) { rlo = llo; rhi = lhi; }       // lanes past the batch: zero-length copies of the last record
        u32 plo = w_shfl_up(rlo, 1), phi = w_shfl_up(rhi, 1);
        if (lane == 0) { plo = carry_lo; phi = carry_hi; }
        const u32 l0 = ((carry_lo >> 28) | ((carry_hi & 0x3FFFu) << 4));
        carry_lo = llo; carry_hi = lhi;
        const bool real = i < nseq;
        const u32 lit_end = (rlo >> 28) | ((rhi & 0x3FFFu) << 4), p_lit_end = (plo >> 28) | ((phi & 0x3FFFu) << 4);
        const u32 out_end = rhi >> 14, p_out_end = phi >> 14;
        const u32 ll = lit_end - p_lit_end;
        const u32 a = p_out_end + sh, mdst = a + ll, bend = out_end + sh, ml = bend - mdst;
        u32 off = rlo & KZG_OFF_MASK;
        if (off & KZG_OFF_SYM) {
            const u32 k = off & 3, dd = (off & (KZG_OFF_SYM - 1)) >> 2;
            const u32 base = k == 0 ? ri0 : (k == 1 ? ri1 : ri2);
            off = (base >= KZG_OFF_INVALID || base <= dd) ? KZG_OFF_INVALID : base - dd;
        }
        // (the sequence kernel guarantees non-decreasing lit_end / out_end within the block limits)
        const bool bad = real && (off - 1u >= KZG_OFF_INVALID - 1u || off > hist_before_block + (mdst - sh));
        const bool usesdict = real && off > frame_before + (mdst - sh);
        const u32 trouble = w_ballot(bad || usesdict);
        if (trouble) status = w_any(bad) ? 1u : 2u;
        const u32 be = w_shfl(bend, last);
        const i32 spos = (i32)mdst - (i32)off;
        const i32 farlim = mx_i(Vlo, (i32)be - (i32)TILE + 256);     // (+256: the next batch's first pass still reads the tail of this batch's last match)
        const bool hasm = real && ml > 0 && !trouble;
        const bool far = hasm && spos < farlim;
        const bool stage = far && ml <= 64u && spos + (i32)ml <= Fdone;
        const u32 mst = sb + L::MST + buf * (32u * KZG_X2_MST) + lane * KZG_X2_MST;
        const u32 a16 = (u32)spos & 15u;
        {
            const u32 nch = stage ? (a16 + ml + 15u) >> 4 : 0u;        // <= 5
            const u8* g = gal + (spos - (i32)a16);
            cp_async16_if(nch > 0, mst, g);
            cp_async16_if(nch > 1, mst + 16, g + 16);
            cp_async16_if(nch > 2, mst + 32, g + 32);
            cp_async16_if(nch > 3, mst + 48, g + 48);
            cp_async16_if(nch > 4, mst + 64, g + 64);
        }
        const bool periodic = hasm && !far && off < 16u && ml > off;
        P.desc = stage ? (KZG_X2_STG | (((mst - sb) + a16 - mdst + KZG_X2_BIAS) & 0xFFFFFFu)) : (off | (periodic ? KZG_X2_PER : 0u));
        P.glob = far && !stage;
        P.a = a; P.ll = ll; P.bend = bend; P.litpos = p_lit_end;
        P.nvalid = nvalid; P.be = be;
        P.has_glob = w_any(P.glob);
        P.has_huge = w_any(ll >= KZG_X2_LLHUGE);
        // ---- literal ring: aligned stream positions (literal index + lalign) [l0, l1) of this batch
        P.litdirect = false;
        if (!lfill) {
            const u32 l1 = (llo >> 28) | ((lhi & 0x3FFFu) << 4);
            const u32 keep = (l0_cur + lalign) & ~15u;                  // the batch that executes next still reads from here on
            const u32 allowed = keep + KZG_X2_LR;
            const u32 want = mn_u(l1, lit_regen) + lalign;
            if (want > allowed) P.litdirect = true;                     // too many literals for the ring: this batch reads them from global memory
            else {
                const u32 from = mx_u(Lf, (l0 + lalign) & ~15u);
                const u32 target = (want + 15u) & ~15u;
                for (u32 c = from + (lane << 4); c < target; c += 512u) {
                    const u32 slot = c & (KZG_X2_LR - 1u);
                    cp_async16_if(true, litring + slot, lit16 + c);
                    cp_async16_if(slot == 0, litring + KZG_X2_LR, lit16 + c);
                }
                Lf = mx_u(from, target);
            }
        }
        cp_commit();
    };

    // ---- execute: the words of batch C, pass by pass
    auto execute = [&](const Batch& C) {
        u32 descP = w_shfl_up(C.desc, 1);
        if (lane == 0) descP = carryP;
        sts128(fld + (lane << 4), make_uint4(C.a | (mn_u(C.ll, KZG_X2_LLHUGE) << 18), C.desc, descP, C.litpos + lalign - C.a));
        sts32(fld2 + (lane << 2), C.a + C.ll);
        const u32 fw = C.a >> 2, cnt = (C.bend >> 2) - fw;
        const u32 WB = w_shfl(fw, 0), WE = C.be >> 2, NW = WE - WB;
        const bool has_glob = C.has_glob || carry_glob;
        u32 carry_own = 0;
        for (u32 chunk0 = 0; chunk0 < NW; chunk0 += KZG_X2_OWN) {
            const i32 CX0 = (i32)((WB + chunk0) << 2);
            if (CX0 - F > (i32)(TILE / 8)) flush(CX0, false);
            w_sync();                                                  // fields written / the previous chunk's map is no longer read
            if (lane < KZG_X2_OWN / 16) sts128(own + (lane << 4), make_uint4(0, 0, 0, 0));
            w_sync();
            {
                const u32 rel = fw - WB - chunk0;
                sts8_if(cnt > 0 && rel < KZG_X2_OWN, own + (rel & (KZG_X2_OWN - 1u)), lane + 1u);
            }
            w_sync();
            const u32 nw_chunk = mn_u(NW - chunk0, KZG_X2_OWN);
            for (u32 p0 = 0; p0 < nw_chunk; p0 += 32) {
                const u32 idx = p0 + lane;
                const bool wvalid = idx < nw_chunk;
                const i32 PX0 = CX0 + (i32)(p0 << 2);
                const i32 X0 = PX0 + (i32)(lane << 2);
                const i32 pass_end = mn_i(PX0 + 128, (i32)(WE << 2));
                const i32 Tv = mx_i(Vlo, pass_end - (i32)TILE);         // ring positions >= Tv are valid throughout this pass
                // ---- owner
                const u32 mk = lds8(own + idx);
                const u32 bm = w_ballot(mk != 0);
                const u32 m = bm & (0xFFFFFFFFu >> (31u - lane));
                const u32 ov = w_shfl(mk, (31u - clz32(m)) & 31u);
                const u32 owner = m ? ov - 1u : carry_own;
                carry_own = w_shfl(owner, 31);
                // ---- descriptor of the owner
                const uint4 f = lds128(fld + (owner << 4));
                const u32 a = f.x & 0x3FFFFu, ll_o = f.x >> 18;
                u32 mdst = a + ll_o;
                if (C.has_huge) { if (ll_o == KZG_X2_LLHUGE) mdst = lds32(fld2 + (owner << 2)); }
                const u32 descM = f.y, dP = f.z;
                const i32 dA = (i32)a - X0, dM = (i32)mdst - X0;
                const u32 nP = (u32)mn_i(mx_i(dA, 0), 4), nL = (u32)mn_i(mx_i(dM, 0), 4);
                const u32 mP = lowmask(nP), mLo = lowmask(nL);
                const bool needP = wvalid && nP > 0, needM = wvalid && nL < 4, needL = wvalid && nL > nP;
                const bool stP = (dP >> 31) != 0, stM = (descM >> 31) != 0;
                const u32 offP = dP & KZG_X2_OFF, offM = descM & KZG_X2_OFF;
                const i32 spP = X0 - (i32)offP, spM = X0 - (i32)offM;       // ring sources (meaningful when not staged)
                const u32 addrP = stP ? sb + (u32)X0 + ((dP & 0xFFFFFFu) - KZG_X2_BIAS) : raddr(spP);
                const u32 addrM = stM ? sb + (u32)X0 + ((descM & 0xFFFFFFu) - KZG_X2_BIAS) : raddr(spM);
                // sources that are not (entirely) in the ring any more and were not staged: global memory
                bool gP = false, gM = false, odd = false;
                if (has_glob) {
                    const bool oP = needP && !stP && spP < Tv, oM = needM && !stM && spM < Tv;
                    gP = oP && spP + 4 <= Tv; gM = oM && spM + 4 <= Tv;
                    odd = (oP && !gP) || (oM && !gM);                    // straddles the ring's lower edge: byte by byte
                    if (w_any(oP || oM)) { if (lane == 0) bulk_wait_all(); Fdone = F; S1 = S2 = S3 = S4 = F; w_sync(); }
                }
                // short-period overlapping match, word entirely inside it: read the period
                const bool perw = !stM && (descM & KZG_X2_PER) && wvalid && nL == 0 && (i32)mdst - (i32)offM >= Tv;
                // ---- literal bytes of this word
                u32 vL = rle;
                if (!lfill) {
                    const u32 lpos = f.w + (u32)X0;                     // aligned stream position of the literal byte at X0
                    if (!C.litdirect) vL = ldu32(litring + (lpos & (KZG_X2_LR - 1u)));
                    else {
                        vL = 0;
                        if (needL) for (u32 k = 0; k < 4; k++) { const u32 q = lpos + k - lalign; if (q < lit_regen) vL |= ldg8_ro(lit + q) << (8 * k); }
                    }
                }
                // ---- rounds
                bool pending = wvalid;
                for (;;) {
                    const u32 pm = w_ballot(pending);
                    if (!pm) break;
                    const u32 fl = ffs32(pm) - 1u;
                    const i32 fr = PX0 + (i32)(fl << 2);                // everything below fr is final
                    const bool rP = !needP || stP || gP || (X0 + (i32)nP - (i32)offP <= fr);
                    const bool rM = !needM || stM || gM || (perw ? (i32)mdst <= fr : (X0 + 4 - (i32)offM <= fr));
                    const bool go = pending && rP && rM && !odd;
                    if (w_any(go)) {
                        u32 vP = 0, vM = 0;
                        if (go && needP) vP = gP ? ldu32_hist(gal, spP) : ldu32(addrP);
                        if (go && needM && !perw) vM = gM ? ldu32_hist(gal, spM) : ldu32(addrM);
                        if (w_any(go && perw)) {
                            if (go && perw) {
                                u32 r = ((u32)X0 - mdst) % offM;
                                const i32 base = (i32)mdst - (i32)offM;
                                vM = 0;
                                for (u32 k = 0; k < 4; k++) { vM |= lds8(raddr(base + (i32)r)) << (8 * k); r = r + 1 == offM ? 0 : r + 1; }
                            }
                        }
                        if (go) {
                            const u32 v = (vP & mP) | (vL & mLo & ~mP) | (vM & ~mLo);
                            const u32 slot = (rb + (u32)X0) & (TILE - 1u);
                            sts32(ring + slot, v);
                            sts32_if(slot < 16u, ring + TILE + slot, v);
                        }
                    }
                    const bool slow = pending && !go && lane == fl;
                    if (w_any(slow)) {
                        // the first pending word can always go: everything below it is final; byte by byte, sources inside
                        // the word itself come from the bytes just made
                        if (has_glob) {      // (only then can a byte come from global memory: everything flushed must have landed)
                            if (lane == 0) bulk_wait_all();
                            Fdone = F; S1 = S2 = S3 = S4 = F;
                            w_sync();
                        }
                        if (slow) {
                            u32 v = 0;
                            for (u32 k = 0; k < 4; k++) {
                                const i32 X = X0 + (i32)k;
                                u32 byte;
                                if (k >= nP && k < nL) byte = (vL >> (8 * k)) & 0xFFu;
                                else {
                                    const u32 d = k < nP ? dP : descM;
                                    if (d >> 31) byte = lds8(sb + (u32)X + ((d & 0xFFFFFFu) - KZG_X2_BIAS));
                                    else {
                                        const i32 sp = X - (i32)(d & KZG_X2_OFF);
                                        if (sp >= X0) byte = (v >> (8 * (u32)(sp - X0))) & 0xFFu;
                                        else if (sp >= Tv) byte = lds8(raddr(sp));
                                        else byte = ldg8_hist(gal + sp);
                                    }
                                }
                                v |= byte << (8 * k);
                            }
                            const u32 slot = (rb + (u32)X0) & (TILE - 1u);
                            sts32(ring + slot, v);
                            sts32_if(slot < 16u, ring + TILE + slot, v);
                        }
                    }
                    pending = pending && !go && !slow;
                    w_sync();
                }
            }
        }
        // ---- what the next batch's first word needs of this batch's last match
        {
            const u32 lastdesc = w_shfl(C.desc, C.nvalid - 1u);
            carry_glob = w_shfl((u32)C.glob, C.nvalid - 1u) != 0;
            if (lastdesc >> 31) {
                const i32 Xt = (i32)(WE << 2);
                const i32 nb = (i32)C.be - Xt;                          // 0..3 bytes of the match lie in the next word
                if ((i32)lane < nb) sts8(scr_tail + lane, lds8(sb + (u32)(Xt + (i32)lane) + ((lastdesc & 0xFFFFFFu) - KZG_X2_BIAS)));
                carryP = KZG_X2_STG | ((L::SCR + 16u - (u32)Xt + KZG_X2_BIAS) & 0xFFFFFFu);
            } else carryP = lastdesc;
        }
        if ((i32)(WE << 2) - F >= KZG_X2_FLUSH_MIN) flush((i32)(WE << 2), false);
    };

    Batch N;
    prepare(0, 0, N);
    for (u32 b0 = 0; b0 < ntot && !status; b0 += 32) {
        const Batch C = N;
        if (b0 + 32 < ntot) prepare(b0 + 32, ((b0 >> 5) + 1u) & 1u, N);     // (the literal ring keeps this batch's literals)
        else cp_commit();
        if (status) break;
        cp_wait<1>();
        w_sync();
        execute(C);
        l0_cur = w_shfl(C.litpos + C.ll, C.nvalid - 1u);       // first literal of the batch that executes next
    }
    cp_wait<0>();
    w_sync();
    if (!status) flush(cur_end, true);
    if (lane == 0) bulk_wait_all();
    w_sync();
    fence_block();
    return status;
}

// One warp per frame (tickets); blocks of a frame in order.  Raw / RLE blocks, literal-only blocks and long literal
// tails were written by kzg_bulk_kernel before.
#if !defined(KZG_SIMT_EMU)
template <u32 TILE>
__global__ void __launch_bounds__(KZG_X2_WARPS * 32, TILE <= 8192 ? 7 : 4) kzg_exec2_kernel(KzgBatch B) {
    extern __shared__ uint4 smem4[];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 sb = (u32)__cvta_generic_to_shared(smem4) + warp * KzgX2<TILE>::BYTES;
    if (B.ctr->overflow & 3u) return;
    for (;;) {
        u32 f = 0;
        if (lane == 0) f = atomicAdd(&B.ctr->ticket_exec, 1u);
        f = __shfl_sync(0xFFFFFFFFu, f, 0);
        if (f >= B.n_frames) break;
        const KzgFrame* fr = &B.frames[f];
        if (fr->status != KZG_OK) continue;
        u8* const fdst = B.dst + B.dst_off[f];
        const u32 nblk = fr->n_blocks, first = fr->first_block;
        u32 r = 0;
        for (u32 k = 0; k < nblk && r == 0; k++) {
            const KzgBlock blk = B.blocks[first + k];
            if (blk.type != 2) continue;
            const KzgCBlock* cb = &B.cblocks[blk.cidx];
            if (cb->n_seq == 0) continue;
            if (cb->bulk_lits) { r = 2; break; }
            r = kzg_exec2_block<TILE>(B, blk, cb, fdst + blk.out_off, sb, lane);
        }
        if (r == 1 && lane == 0) B.frames[f].status = KZG_E_CORRUPTION;
        if (r == 2 && lane == 0) B.dlist[atomicAdd(&B.ctr->n_defer, 1u)] = f;      // the first design's full variant redoes this frame
    }
}
#endif
