/* TEST INFRASTRUCTURE ONLY.  CPU restatement of the Zstandard decoder that the reference's hot
 * path reaches through ZSTD_decompressStream (reference:
 * kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:178 and
 * kompressor-zstd--nativelib/src/nativeMain/kotlin/com/ensody/kompressor/zstd/ZstdDecompressor.native.kt:62-66).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * The product path (kompressor_b200/csrc) never includes, links or calls anything from oracle/.
 */
#ifndef KZO_ZSTD_ORACLE_H
#define KZO_ZSTD_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* libzstd's ZSTD_ErrorCode values (zstd_errors.h of 1.5.7; names checked against
 * ZSTD_getErrorName of the real library in tests/test_oracle.py). */
enum {
    KZO_OK = 0,
    KZO_E_GENERIC = 1,
    KZO_E_PREFIX_UNKNOWN = 10,
    KZO_E_FRAMEPARAM_UNSUPPORTED = 14,
    KZO_E_WINDOW_TOO_LARGE = 16,
    KZO_E_CORRUPTION = 20,
    KZO_E_CHECKSUM_WRONG = 22,
    KZO_E_LITERALS_HEADER_WRONG = 24,
    KZO_E_DICT_CORRUPTED = 30,
    KZO_E_DICT_WRONG = 32,
    KZO_E_TABLELOG_TOO_LARGE = 44,
    KZO_E_MAXSYMBOL_TOO_LARGE = 46,
    KZO_E_DSTSIZE_TOO_SMALL = 70,
    KZO_E_SRCSIZE_WRONG = 72
};

typedef struct {
    uint32_t ll;     /* literal length */
    uint32_t ml;     /* match length */
    uint32_t offset; /* resolved distance back (repeat codes already applied) */
} kzo_seq;

/* One decoded block, for stage-by-stage comparison with the CUDA kernels. */
typedef struct {
    int type;            /* 0 raw, 1 rle, 2 compressed */
    int last;
    uint32_t content_size;  /* Block_Size field */
    uint32_t regen_size;    /* bytes this block produced */
    int lit_type;           /* 0 raw 1 rle 2 compressed 3 treeless; -1 if not compressed */
    uint32_t lit_size;
    int n_streams;
    uint32_t n_seq;
    int modes[3];           /* LL, OF, ML */
    int logs[3];            /* accuracy logs in use, LL, OF, ML */
    int huf_maxbits;
} kzo_block_info;

typedef struct {
    uint64_t window_size;
    uint64_t content_size;   /* UINT64_MAX if absent */
    uint32_t dict_id;
    int single_segment;
    int has_checksum;
    uint32_t header_size;
} kzo_frame_header;

/* Callback invoked after each block is decoded (may be NULL). literals/seqs are valid only
 * during the call. */
typedef void (*kzo_block_cb)(void* user, const kzo_block_info* info, const uint8_t* literals,
                             const kzo_seq* seqs);

/* Decodes every frame in [src, src+src_size) (concatenated frames, skippable frames allowed)
 * into dst.  Returns the number of bytes written, or (size_t)-code on error, exactly like
 * libzstd's size_t convention.  dict may be NULL (raw-content or formatted dictionary). */
size_t kzo_decompress(const uint8_t* src, size_t src_size, uint8_t* dst, size_t dst_cap,
                      const uint8_t* dict, size_t dict_size);

size_t kzo_decompress_ex(const uint8_t* src, size_t src_size, uint8_t* dst, size_t dst_cap,
                         const uint8_t* dict, size_t dict_size, kzo_block_cb cb, void* user);

/* Parses one frame header. Returns 0, or (size_t)-code; (size_t)-KZO_E_SRCSIZE_WRONG if the
 * header is incomplete. */
size_t kzo_frame_header_parse(const uint8_t* src, size_t src_size, kzo_frame_header* out);

/* Size in bytes of the first frame (zstd or skippable) in src, walking block headers only.
 * Returns (size_t)-code on error / truncation. */
size_t kzo_find_frame_size(const uint8_t* src, size_t src_size);

uint64_t kzo_xxh64(const uint8_t* p, size_t len, uint64_t seed);

int kzo_is_error(size_t code);
const char* kzo_error_name(size_t code);

#ifdef __cplusplus
}
#endif
/* 1 if a frame decoded since the last call ended because a Huffman stream ran out of bits where libzstd's fast loops
   make up literals (the documented divergence); clears the flag.  Not thread safe (test tooling). */
int kzo_made_up_literals(void);

#endif
