// kzg = "Kompressor zstd on GPU": B200-native Zstandard decoder behind Kompressor's SliceTransform API.
// Shared types and small host+device helpers.  Everything marked KZG_HD is plain serial code that
// one GPU thread runs; it also compiles as host C++ so tests/hostsim can exercise the exact same
// logic on a CPU (test infrastructure only -- the product library never runs it on the host).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define KZG_HD __host__ __device__ __forceinline__
#define KZG_D __device__ __forceinline__
#else
#define KZG_HD inline
#endif

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int16_t i16;
typedef int32_t i32;
typedef int64_t i64;

// libzstd ZSTD_ErrorCode numbering (zstd_errors.h 1.5.7); what Kompressor's checkErrorResult sees
// (reference: kompressor-zstd--nativelib/src/jvmCommonMain/kotlin/com/ensody/kompressor/zstd/ZstdCompressor.jvm.kt:45-51).
enum : u32 {
    KZG_OK = 0,
    KZG_E_GENERIC = 1,
    KZG_E_PREFIX_UNKNOWN = 10,
    KZG_E_FRAMEPARAM_UNSUPPORTED = 14,
    KZG_E_WINDOW_TOO_LARGE = 16,
    KZG_E_CORRUPTION = 20,
    KZG_E_CHECKSUM_WRONG = 22,
    KZG_E_LITERALS_HEADER_WRONG = 24,
    KZG_E_DICT_CORRUPTED = 30,
    KZG_E_DICT_WRONG = 32,
    KZG_E_MEMORY = 64,
    KZG_E_DSTSIZE_TOO_SMALL = 70,
    KZG_E_SRCSIZE_WRONG = 72,
    KZG_E_INTERNAL_OVERFLOW = 119   // workspace too small: host grows it and re-runs (never surfaced)
};

#define KZG_BLOCK_MAX (128u * 1024u)
#define KZG_BULK_MIN 256u            // literal runs of at least this many bytes are copied by the bulk kernel ...
#define KZG_BULK_SUM 4096u           // ... in blocks whose long runs add up to at least this much (else not worth a record scan)
#define KZG_HUF_MAXBITS 11
#define KZG_MAGIC 0xFD2FB528u
#define KZG_SKIP_MAGIC 0x184D2A50u
#define KZG_DICT_MAGIC 0xEC30A437u
#define KZG_NONE 0xFFFFFFFFu

// Sequence record: off_enc[0,28) | lit_end[28,46) | out_end[46,64).
//  lit_end / out_end are running totals inside the block *after* this sequence, so the execute
//  kernel needs no prefix sum.  off_enc is the resolved match distance, or (bit 27 set) a symbolic
//  reference "initial repeat offset k minus d" for blocks whose starting repeat history is not
//  known when they are entropy-decoded in parallel: k = bits[0,2), d = bits[2,27).
#define KZG_OFF_SYM 0x8000000u
#define KZG_OFF_MASK 0xFFFFFFFu
#define KZG_OFF_INVALID 0x7FFFFFFu   // concrete but impossible distance -> corruption at execution

KZG_HD u64 kzg_pack_seq(u32 off_enc, u32 lit_end, u32 out_end) {
    return (u64)off_enc | ((u64)lit_end << 28) | ((u64)out_end << 46);
}

// ---------------------------------------------------------------------------------------------
// Work descriptors in the device workspace.
struct KzgFrame {           // one per input frame
    u64 window_size;
    u64 fcs;                // Frame_Content_Size or ~0
    u64 out_size;           // bytes produced (valid when status == 0)
    u32 first_block;        // index into KzgBlock[]
    u32 n_blocks;
    u32 hdr_size;
    u32 dict_id;
    u32 checksum_off;       // offset of the 4-byte XXH64 trailer inside the frame, or KZG_NONE
    u32 status;             // frame-level error found by the scan
    u32 block_max;
    u32 xxh_lo;             // low 32 bits of XXH64 computed by the checksum kernel
    u32 lit_garbage;        // a block's Huffman stream ran out of bits where libzstd's fast loops make up literals (see kzg_huf_verdict)
    u32 fail_block;         // status != 0: index (in the frame) of the block the error belongs to; the blocks before it are good
    // entropy state after the last block, for a frame that is decoded in segments (kzg_decompress_segment_host):
    u32 end_owner[4];       // table owners: Huffman, LL, OF, ML (a compressed block, KZG_NONE, KZG_DICT_OWNER or KZG_INVALID_OWNER)
    u32 end_rep[3];         // repeat offsets
    u32 pad_;
};

struct KzgBlock {           // one per block of every frame
    u64 src;                // absolute offset of the block content in the source buffer
    u64 out_off;            // offset of this block's output inside its frame's output
    u32 frame;
    u32 csize;              // Block_Size field
    u32 regen;              // regenerated size (raw/RLE: known at scan; compressed: after sequence decode)
    u32 cidx;               // index into KzgCBlock[] or KZG_NONE
    u32 type;               // 0 raw, 1 RLE, 2 compressed
    u32 last;
};

struct KzgCBlock {          // one per compressed block
    u64 src;                // absolute offset of block content
    u64 lit_out;            // byte offset in the literal pool (Huffman-coded literals only)
    u64 seq_out;            // record offset in the sequence pool
    u32 block;              // index into KzgBlock[]
    u32 frame;
    u32 csize;
    u32 status_lit;         // first error found while parsing/decoding literals
    u32 status_seq;         // ... while parsing/decoding sequences
    u32 status_exec;        // ... while executing
    // literals
    u32 lit_type;           // 0 raw, 1 RLE, 2 Huffman, 3 treeless
    u32 lit_regen;
    u32 lit_src;            // offset from block start: raw bytes / RLE byte / first Huffman stream
    u32 n_streams;          // 0, 1 or 4
    u32 stream_size[4];
    u32 huf_owner;          // cblock whose weights describe the table (self, earlier block, or KZG_NONE)
    // sequences
    u32 n_seq;
    u32 modes;              // LL | OF<<2 | ML<<4  (0 predefined, 1 RLE, 2 FSE, 3 repeat)
    u32 fse_owner[3];       // cblock whose description to use for LL, OF, ML (KZG_NONE = predefined)
    u32 seq_src;            // offset from block start of the sequence bitstream
    u32 seq_size;           // its size in bytes
    u32 regen;              // total output of this block
    u32 lit_used;           // literals consumed by sequences
    u32 rep_init[3];        // concrete repeat offsets at block start (filled by the frame-order pass)
    u32 rep_final[3];       // encoded (possibly symbolic) repeat offsets after the block
    u32 first_in_frame;     // 1 if no compressed block with sequences precedes it in its frame
    u32 bulk_lits;          // 1 if the literal runs of >= KZG_BULK_MIN bytes add up to >= KZG_BULK_SUM: the bulk kernel copies them
    u32 lit_garbage;        // a Huffman stream ran out of bits in a block libzstd decodes with its fast loops
};

struct KzgFseDesc {         // entropy description parsed from a compressed block
    i16 norm[3][64];        // LL (36), OF (32), ML (53) normalised counts
    u8 log[3];              // accuracy logs; for RLE mode log = 0 and norm[t][0] = symbol
    u8 nsym[3];
    u8 kind[3];             // 0 none, 1 RLE, 2 FSE
    u8 pad[7];              // sizeof == 400: every norm[t] row stays 16-byte aligned in an array
};

struct KzgHufDesc {
    u8 weight[256];
    u32 nsym;               // number of weights including the implicit last one
    u32 maxbits;
    u32 pad[2];             // sizeof == 272 (16-byte multiple)
};

struct KzgCounters {
    u32 n_blocks, n_cblocks, n_hblocks, n_sblocks;
    u64 lit_used, seq_used;
    u32 overflow;           // bit0 blocks, bit1 cblocks, bit2 literal pool, bit3 sequence pool
    u32 ticket_huf, ticket_fse, ticket_exec, ticket_raw;
    u32 n_bulk;             // entries of blist
    u32 n_defer;            // entries of dlist
    u32 ticket_exec2;
    u32 n_hblocks2, ticket_huf2;   // Huffman blocks whose table does not fit the compact two-level layout (second launch)
    u32 ticket_exec4;              // the CTA-per-frame execute kernel's ticket over dlist
};

#define KZG_DEFER_BIG 0x80000000u
#define KZG_INVALID_OWNER 0xFFFFFFFEu
#define KZG_DICT_OWNER 0xFFFFFFFDu      // table comes from the formatted dictionary loaded into the context

struct KzgBatch {
    // caller's frame table (device pointers)
    const u8* src;
    const u64* src_off;
    const u64* src_size;
    u8* dst;
    const u64* dst_off;
    const u64* dst_cap;
    u32* status;        // out: ZSTD_ErrorCode per frame (0 = ok)
    u64* out_size;      // out: decoded bytes per frame
    u32 n_frames;
    u32 tile_bytes;     // shared-memory output tile of the execute kernel (>= largest block)
    u32 big_by_cta;     // frames of many blocks are left to the CTA-per-frame execute kernel (set at launch)
    // segment mode (a frame decoded piece by piece behind the streaming shim): the batch is one synthetic frame that
    // continues a real one -- its tables start from init_owner[] (Huffman, LL, OF, ML), its repeat offsets from dict_rep[],
    // its history is `dict`, and the block size limit is the real frame's
    u32 seg_mode;
    u32 init_owner[4];
    u32 seg_block_max;
    u32 huf_long_min;   // blocks with at least this many Huffman-coded literals go to the full-table launch (set at launch)
    u32 lit_ride;       // literal warps in the sequence kernel: 0 = where they can finish (estimate), 1 = always, 2 = never (set at launch)
    // workspace
    KzgFrame* frames;
    KzgBlock* blocks;
    KzgCBlock* cblocks;
    KzgFseDesc* fse;
    KzgHufDesc* huf;
    u32* hlist;         // compressed blocks with Huffman-coded literals
    u32* hlist2;        // ... of those, the ones left to the second Huffman launch (worst-case table layout)
    u32* slist;         // compressed blocks with sequences
    u32* blist;         // blocks with work for the bulk kernel (filled by the offsets stage)
    u32* dlist;         // frames the hot execute kernel left to the others: long runs (full variant); KZG_DEFER_BIG set: many blocks (CTA per frame)
    u8* lit_pool;
    u64* seq_pool;
    u16* fse_tab;       // built FSE decode tables, 1280 cells (LL 512 | ML 512 | OF 256) per entry of slist
    KzgCounters* ctr;
    const KzgFseDesc* fse_default;
    const u32* code_tables;   // 36 LL + 53 ML packed (extra_bits << 24 | baseline)
    u64 lit_cap, seq_cap;
    u32 block_cap, cblock_cap;
    // optional dictionary shared by all frames of the batch: `dict` is its content (history prefix); a
    // formatted dictionary (magic 0xEC30A437) also brings entropy tables and repeat offsets
    const u8* dict;
    u32 dict_size;
    u32 dict_id;
    u32 dict_entropy;           // 1 if fse_dict / huf_dict / dict_rep are valid
    u32 dict_rep[3];
    const KzgFseDesc* fse_dict;
    const KzgHufDesc* huf_dict;
};


// ---------------------------------------------------------------------------------------------
KZG_HD u32 kzg_rd16(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8); }
KZG_HD u32 kzg_rd24(const u8* p) { return kzg_rd16(p) | ((u32)p[2] << 16); }
KZG_HD u32 kzg_rd32(const u8* p) { return kzg_rd16(p) | (kzg_rd16(p + 2) << 16); }
KZG_HD u64 kzg_rd64(const u8* p) { return (u64)kzg_rd32(p) | ((u64)kzg_rd32(p + 4) << 32); }

KZG_HD int kzg_highbit(u32 v) {   // floor(log2(v)), v > 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

KZG_HD u32 kzg_funnel_r(u32 lo, u32 hi, u32 sh) {   // low 32 bits of (hi:lo) >> sh, 0 <= sh < 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

static_assert(sizeof(KzgFseDesc) == 400, "KzgFseDesc rows must stay 16-byte aligned");
static_assert(sizeof(KzgHufDesc) == 272, "KzgHufDesc must be a 16-byte multiple");
