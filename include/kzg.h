/* kzg.h -- C ABI of the B200-native Zstandard decompressor ("Kompressor zstd on GPU").
 *
 * This is the drop-in boundary for Kompressor's zstd decode path.  Plain C: pointers and sizes only.
 * Three layers:
 *
 *  Layer A  batch API -- the measured path.  Independent frames decoded as one GPU batch.
 *  Layer B  streaming shim with libzstd semantics under the kzg_ prefix.  One-for-one replacements of
 *           the six libzstd functions the reference binds for decompression:
 *             ZSTD_createDCtx           <- Wrapper.cpp:128, ZstdDecompressor.native.kt:32
 *             ZSTD_freeDCtx             <- Wrapper.cpp:139, ZstdDecompressor.native.kt:37
 *             ZSTD_DCtx_loadDictionary  <- Wrapper.cpp:70,  ZstdDecompressor.native.kt:42
 *             ZSTD_decompressStream     <- Wrapper.cpp:178, ZstdDecompressor.native.kt:62-66
 *             ZSTD_isError              <- Wrapper.cpp:192, ZstdDecompressor.native.kt:69
 *             ZSTD_getErrorName         <- Wrapper.cpp:195, ZstdDecompressor.native.kt:70
 *           (paths relative to kompressor-zstd--nativelib/src/{jvmCommonMain/jni,nativeMain/kotlin/...}).
 *           The kzg_ prefix is required because the compressor side keeps using the real libzstd in
 *           the same process.
 *  Layer C  JNI symbols Java_com_ensody_kompressor_zstd_ZstdWrapper_* (kzg_jni.cpp), compiled only
 *           where <jni.h> exists; see INTEGRATION.md.
 *
 * There is no CPU fallback: every entry point that decodes requires a CUDA device and fails with
 * an error if none is usable.
 */
#ifndef KZG_H
#define KZG_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KZG_VERSION_STRING "0.1.0"

/* Error convention: identical to libzstd's.  Functions returning size_t return (size_t)-code on
 * error; kzg_isError(r) is r > (size_t)-120.  Codes are libzstd 1.5.7's ZSTD_ErrorCode values, which
 * is what Kompressor's checkErrorResult prints
 * (kompressor-zstd--nativelib/src/jvmCommonMain/kotlin/com/ensody/kompressor/zstd/ZstdCompressor.jvm.kt:45-51). */
unsigned kzg_isError(size_t code);
const char* kzg_getErrorName(size_t code);
const char* kzg_versionString(void);

/* ------------------------------------------------------------------ Layer A: batch API */
typedef struct kzg_ctx kzg_ctx;

/* Creates a decode context on CUDA device `device` (-1 = current device). Returns 0 on success,
 * otherwise a positive ZSTD_ErrorCode (64 = allocation / no usable device). */
int kzg_ctx_create(kzg_ctx** out, int device);
void kzg_ctx_destroy(kzg_ctx* ctx);

/* Frame table builder (host). Walks concatenated zstd / skippable frames in [buf, buf+len) using only
 * frame and block headers and records up to max_frames entries:
 *   off[i], size[i]   position and compressed size of frame i
 *   content[i]        Frame_Content_Size if present, else an upper bound (n_blocks * block_max);
 *                     0 for skippable frames
 * *consumed = bytes covered by complete frames.  Stops without error at a truncated frame.
 * Returns the number of frames, or (size_t)-code (for example unknown magic -> -10). */
size_t kzg_scan_frames(const void* buf, size_t len, uint64_t* off, uint64_t* size, uint64_t* content,
                       size_t max_frames, size_t* consumed);

/* Seek table of the zstd seekable format: a skippable frame {0x184D2A5E, size, n x {compressed size, content size},
 * n, descriptor, 0x8F92EAB1} appended to a stream of frames ("independent frames with
 * a seekable index").  kzg_write_seek_table writes it for n frames (returns its size, kzg_seek_table_bound(n), or an
 * error code).  kzg_scan_frames uses a trailing table when it describes the buffer exactly (one pass over 8 bytes per
 * frame instead of a walk over every block header) and falls back to the walk otherwise; the table's own skippable
 * frame is reported as the last entry, like any skippable frame. */
size_t kzg_seek_table_bound(size_t n_frames);
size_t kzg_write_seek_table(void* dst, size_t cap, const uint64_t* comp_size, const uint64_t* content_size, size_t n_frames);

/* Decodes n_frames independent frames whose bytes are already in device memory.
 *   d_src + d_src_off[i], d_src_size[i]   exactly one zstd (or skippable) frame each
 *   d_dst + d_dst_off[i], d_dst_cap[i]    where its output goes and how much room there is
 *   d_status[i]                           out: 0 or ZSTD_ErrorCode (20 corruption, 22 checksum, 70 dst too small...)
 *   d_out_size[i]                         out: decoded bytes.  For a frame that failed: the index of the block the error
 *                                         belongs to (the blocks before it are good: the streaming layer hands them over
 *                                         first, as libzstd does); for status 22 the decoded size -- the content is in place,
 *                                         its checksum does not match -- unless bit 63 is set (then: block index in the low
 *                                         30 bits; bit 62 marks the kind of error DESIGN.md "Error behaviour" documents)
 * All pointers are device pointers; the 16-byte aligned chunks around every byte of d_src must be
 * readable (true for any cudaMalloc allocation).  total_src_bytes / total_dst_cap / max_frame_dst_cap
 * are host-side sums / max of the table, used to size the workspace.  Work is enqueued on
 * `cuda_stream` (a cudaStream_t, NULL = default stream) and NOT synchronised.
 * Returns the number of kernels launched (> 0), or a negative ZSTD_ErrorCode. */
int kzg_decompress_batch_device(kzg_ctx* ctx, const void* d_src, const uint64_t* d_src_off,
                                const uint64_t* d_src_size, void* d_dst, const uint64_t* d_dst_off,
                                const uint64_t* d_dst_cap, uint64_t n_frames, uint64_t total_src_bytes,
                                uint64_t total_dst_cap, uint64_t max_frame_dst_cap, uint32_t* d_status,
                                uint64_t* d_out_size, void* cuda_stream);

/* Synchronises the stream used by the last kzg_decompress_batch_device call and reports whether the
 * workspace was too small (optimistic sizing).  Returns 0 = results valid, 1 = workspace has been
 * grown: enqueue the same batch again, negative = CUDA failure. */
int kzg_batch_finish(kzg_ctx* ctx);

/* Same batch, host memory: copies the frames to the device, decodes, copies output, status and
 * sizes back (synchronous; retries internally if the workspace had to grow).  Pinned host memory
 * (kzg_host_alloc) makes the copies asynchronous and full-speed.  Returns 0 or a negative code. */
int kzg_decompress_batch_host(kzg_ctx* ctx, const void* src, const uint64_t* src_off, const uint64_t* src_size,
                              void* dst, const uint64_t* dst_off, const uint64_t* dst_cap, uint64_t n_frames,
                              uint32_t* status, uint64_t* out_size);

void* kzg_host_alloc(size_t bytes);   /* pinned host memory */
void kzg_host_free(void* p);

/* Statistics of the last batch (for bench.py): kernels launched, workspace bytes, device time is
 * measured by the caller with CUDA events. */
typedef struct {
    uint32_t n_blocks, n_cblocks, n_hblocks, n_sblocks;
    uint64_t lit_bytes, seq_records, workspace_bytes;
    uint32_t launches, tile_bytes, sm_count, overflow;
} kzg_batch_stats;
int kzg_last_batch_stats(kzg_ctx* ctx, kzg_batch_stats* out);

/* Dictionary used by the following batches (NULL clears it): a raw-content dictionary is a history prefix;
 * a formatted one (magic 0xEC30A437) also preloads every frame's Huffman / FSE tables and repeat offsets and
 * carries a dictionary ID that frames naming another ID are rejected against (dictionary_wrong, 32).
 * Replaces ZSTD_DCtx_loadDictionary (Wrapper.cpp:58-73) on the batch layer. */
int kzg_ctx_set_dictionary(kzg_ctx* ctx, const void* dict, size_t size);

/* ---- A frame decoded in SEGMENTS (what lets the streaming shim below hand over output block by block, as
 * ZSTD_decompressStream does, with memory bounded by the window): a segment is a run of whole blocks of one frame, in
 * order.  The state carries what the next segment needs from the previous ones -- the entropy tables that "repeat" and
 * "treeless" modes refer to, the three repeat offsets, and the running XXH64 of the content.  The caller keeps the last
 * `window` bytes of output and passes them as `history`.
 *   kzg_segment_begin   state at the start of a frame (tables / offsets / history prefix come from `dict`, if any);
 *                       returns the number of dictionary content bytes (the initial history is the end of `dict`)
 *   kzg_decompress_segment_host   blocks: the segment's blocks with their 3-byte headers; last_hdr_off: offset of the
 *                       last block's header in it (its Last_Block bit is forced on in a private copy);
 *                       *status / *out_size as in the batch calls; the state is advanced only when *status == 0
 *   kzg_segment_checksum          low 32 bits of the XXH64 of everything decoded since kzg_segment_begin */
typedef struct kzg_seg_state {
    uint32_t owner_kind[4];       /* Huffman, LL, OF, ML tables: 0 none yet, 1 predefined, 2 the description below */
    uint32_t rep[3];
    uint32_t block_max;           /* Block_Maximum_Size of the frame */
    uint64_t xxh_acc[4], xxh_total;
    uint8_t xxh_tail[32];         /* content bytes not yet folded into the accumulators (xxh_total % 32 of them) */
    uint32_t want_checksum, digest;   /* digest: low 32 bits of the XXH64 of the content so far (kept current by every segment) */
    uint64_t fse_desc[50];        /* KzgFseDesc */
    uint64_t huf_desc[34];        /* KzgHufDesc */
} kzg_seg_state;
long long kzg_segment_begin(kzg_seg_state* st, uint32_t block_max, int want_checksum, const void* dict, size_t dict_size);
int kzg_decompress_segment_host(kzg_ctx* ctx, kzg_seg_state* st, const void* blocks, size_t blocks_size, size_t last_hdr_off,
                                const void* history, size_t history_size, void* dst, size_t dst_cap,
                                uint32_t* status, uint64_t* out_size);
uint32_t kzg_segment_checksum(const kzg_seg_state* st);

/* Per-stage device timing with CUDA events recorded on the launching stream (bench.py's roofline).
 * kzg_last_stage_ms fills ms[i] / names[i] for the stages of the last finished batch and returns
 * how many stages it wrote (0 if profiling was off). */
int kzg_ctx_set_profiling(kzg_ctx* ctx, int enable);
int kzg_last_stage_ms(kzg_ctx* ctx, float* ms, const char** names, int max_stages);
/* The same for the batch enqueued `age` batches before the newest (0 = the newest; the last 32 profiled batches are
 * kept), so that batches enqueued back to back need no host synchronisation in between to be timed.  The caller
 * synchronises the stream first.  Returns 0 if that batch is not (or no longer) kept. */
int kzg_stage_ms_history(kzg_ctx* ctx, int age, float* ms, const char** names, int max_stages);

/* ------------------------------------------------------------------ Layer A, compress side: batch frame writer
 * SURVEY.md section 8(f) rank 4, the step before the path.  Replaces, for producers that hold many chunks,
 * ZSTD_compressStream2 behind Java_..._ZstdWrapper_compressStream (Wrapper.cpp:75-121, ZstdWrapper.kt:26-46):
 * every chunk (at most 128 KiB; 64 KiB is the shape the decoder is measured on) becomes one independent
 * single-block zstd frame with its content size in the header, i.e. exactly what layer A decodes in bulk.
 * Any conforming decoder (libzstd included) reads the output. */
typedef struct kzg_enc kzg_enc;
int kzg_enc_create(kzg_enc** out, int device);                 /* device < 0: current device */
void kzg_enc_destroy(kzg_enc* enc);
size_t kzg_compress_bound(size_t src_size);                    /* capacity every chunk's frame needs (ZSTD_compressBound's role) */
#define KZG_COMPRESS_CHECKSUM 1                                /* flags: append the XXH64 content checksum (ZSTD_c_checksumFlag) */
#define KZG_COMPRESS_LAZY 2                                    /* flags: one step of lazy matching: ~2 % smaller output, match stage ~25 % slower */

/* Compresses n chunks resident in device memory: chunk i is d_src + d_src_off[i], d_src_size[i] bytes; its frame
 * goes to d_dst + d_dst_off[i] (capacity d_dst_cap[i] >= kzg_compress_bound(size), else status 70).
 * d_status[i]: 0 or ZSTD_ErrorCode (72 = chunk larger than 128 KiB); d_out_size[i]: frame bytes.
 * max_chunk = largest d_src_size (host-side, sizes the workspace).  The 16-byte aligned chunks around every byte
 * of d_src must be readable.  Enqueued on `cuda_stream` (NULL: the encoder's own), not synchronised: call
 * kzg_enc_finish.  Returns the number of kernels launched or a negative code. */
int kzg_compress_batch_device(kzg_enc* enc, const void* d_src, const uint64_t* d_src_off, const uint64_t* d_src_size,
                              void* d_dst, const uint64_t* d_dst_off, const uint64_t* d_dst_cap, uint64_t n,
                              uint64_t max_chunk, int flags, uint32_t* d_status, uint64_t* d_out_size, void* cuda_stream);
int kzg_enc_finish(kzg_enc* enc, void* cuda_stream);
/* Same with host memory (synchronous). */
int kzg_compress_batch_host(kzg_enc* enc, const void* src, const uint64_t* src_off, const uint64_t* src_size,
                            void* dst, const uint64_t* dst_off, const uint64_t* dst_cap, uint64_t n, int flags,
                            uint32_t* status, uint64_t* out_size);
int kzg_enc_set_profiling(kzg_enc* enc, int enable);
int kzg_enc_last_stage_ms(kzg_enc* enc, float* ms, const char** names, int max_stages);

/* ------------------------------------------------------------------ Layer B: streaming shim */
typedef struct kzg_DCtx kzg_DCtx;
typedef struct { const void* src; size_t size; size_t pos; } kzg_inBuffer;   /* == ZSTD_inBuffer */
typedef struct { void* dst; size_t size; size_t pos; } kzg_outBuffer;        /* == ZSTD_outBuffer */

kzg_DCtx* kzg_createDCtx(void);                                               /* NULL if no GPU */
size_t kzg_freeDCtx(kzg_DCtx* dctx);
size_t kzg_DCtx_loadDictionary(kzg_DCtx* dctx, const void* dict, size_t dictSize);
/* Returns 0 when a frame is completely decoded and flushed, a positive hint when more input or
 * output room is needed, or an error.  Always consumes all offered input (Kompressor never sets
 * input.insufficient for zstd) and keeps returning non-zero while it still holds undelivered output,
 * which is what `output.insufficient = output.isFull && result != 0` needs
 * (ZstdDecompressor.jvm.kt:38, ZstdDecompressor.native.kt:72). */
size_t kzg_decompressStream(kzg_DCtx* dctx, kzg_outBuffer* output, kzg_inBuffer* input);

/* Compress side of the shim, on top of the batch frame writer: the libzstd functions
 * behind ZstdWrapper.createCompressor / freeCompressor / setParameter / loadCompressorDictionary / compressStream
 * (Wrapper.cpp:10-56,75-121; ZstdCompressor.jvm.kt:10-42, ZstdCompressor.native.kt:28-80).
 *   kzg_createCCtx            ZSTD_createCCtx (NULL if no GPU)
 *   kzg_CCtx_setParameter     ZSTD_CCtx_setParameter: 201 (checksumFlag) is honoured; 100 (compressionLevel) selects the
 *                             parse (below 3: greedy, from 3: one step of lazy matching); 101-107, 200, 202 are accepted
 *                             (content size always written); others: parameter_unsupported (40)
 *   kzg_CCtx_loadDictionary   raw-content dictionaries are accepted (and not referenced); formatted ones: 40
 *   kzg_compressStream2       ZSTD_compressStream2; end_op 0 = ZSTD_e_continue, 1 = flush, 2 = ZSTD_e_end.  All input
 *                             is consumed; the result is the number of bytes still to be flushed (0 = done), which
 *                             is what `output.insufficient = input.hasData || (finish && result != 0)` needs.
 * The stream is written as independent 64 KiB frames (not one long frame): any zstd decoder reads the
 * concatenation, and the batch decoder reads it in bulk. */
typedef struct kzg_CCtx kzg_CCtx;
kzg_CCtx* kzg_createCCtx(void);
size_t kzg_freeCCtx(kzg_CCtx* cctx);
size_t kzg_CCtx_setParameter(kzg_CCtx* cctx, int param, int value);
size_t kzg_CCtx_loadDictionary(kzg_CCtx* cctx, const void* dict, size_t dictSize);
size_t kzg_compressStream2(kzg_CCtx* cctx, kzg_outBuffer* output, kzg_inBuffer* input, int end_op);

#ifdef __cplusplus
}
#endif
#endif
