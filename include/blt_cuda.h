/*
 * blt_cuda.h -- C ABI of libblt_cuda.so: the B200 (sm_100a) implementation of blt_core's
 * tokenization hot path (jtrefon/blt v0.2.2).  Plain C, plain pointers and sizes; no C++ or torch
 * types cross this boundary.  Each entry point names the reference interface it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the Rust-side binding.
 *
 * Conventions
 *   - every function returning int returns BLT_OK (0) or a negative blt_status; the message for
 *     the calling thread's last failure is blt_last_error().
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails with
 *     BLT_ERR_NO_DEVICE / BLT_ERR_CUDA.
 *   - handles are opaque, created and destroyed by the caller.  One blt_strategy may be used
 *     from many host threads at once (reference: `TokenizationStrategy: Send + Sync`,
 *     blt_core/src/tokenizer.rs:21, up to num_threads concurrent calls, pipeline.rs:86-97).
 *   - token streams are big-endian u16 without framing (tokenizer.rs:88-91, 116-120).
 */
#ifndef BLT_CUDA_H
#define BLT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BLT_API __attribute__((visibility("default")))
#else
#define BLT_API
#endif

/* Status codes.  The first four mirror the std::io::ErrorKind values the reference produces. */
typedef enum blt_status {
    BLT_OK = 0,
    BLT_ERR_NOT_FOUND = -1,     /* io::ErrorKind::NotFound     (File::open: config_loader.rs:15, io_handler.rs:54) */
    BLT_ERR_INVALID_INPUT = -2, /* io::ErrorKind::InvalidInput (lib.rs:176-182, lib.rs:194-201)       */
    BLT_ERR_INVALID_DATA = -3,  /* io::ErrorKind::InvalidData  (config_loader.rs:27-43)               */
    BLT_ERR_IO = -4,            /* any other io::Error                                                  */
    BLT_ERR_CUDA = -5,          /* a CUDA runtime call or kernel failed                                 */
    BLT_ERR_NOMEM = -6,         /* host or device allocation failed                                     */
    BLT_ERR_CAPACITY = -7,      /* caller's output buffer is too small                                  */
    BLT_ERR_NO_DEVICE = -8      /* no CUDA device / driver: the product has no CPU path                 */
} blt_status;

/* ContentType (blt_core/src/lib.rs:81-104).  BLT_CONTENT_NONE = Option::None. */
typedef enum blt_content_type {
    BLT_CONTENT_NONE = -1,
    BLT_CONTENT_TEXT = 0,  /* 0xFF01 */
    BLT_CONTENT_AUDIO = 1, /* 0xFF02 */
    BLT_CONTENT_BIN = 2,   /* 0xFF03 */
    BLT_CONTENT_VIDEO = 3  /* 0xFF04 */
} blt_content_type;

typedef struct blt_ctx blt_ctx;           /* one CUDA device: streams, pinned staging, workspaces */
typedef struct blt_strategy blt_strategy; /* Arc<dyn TokenizationStrategy> (tokenizer.rs:21-31)   */

/* ---- library / device ------------------------------------------------------------------------ */

/* env!("CARGO_PKG_VERSION") as returned by blt_python `version()` (blt_python/src/lib.rs:205-208). */
BLT_API const char *blt_version(void);
/* Message of the last failure on the calling thread ("" if none). Never NULL. */
BLT_API const char *blt_last_error(void);
/* Number of usable CUDA devices; BLT_ERR_NO_DEVICE if there is no driver/device. */
BLT_API int blt_device_count(int *count);

BLT_API int blt_ctx_create(int device, blt_ctx **out);
BLT_API void blt_ctx_destroy(blt_ctx *ctx);

/* ---- strategies: select_strategy (blt_core/src/lib.rs:271-282) ---------------------------------- */

/* BasicTokenizationStrategy (tokenizer.rs:103-124). */
BLT_API int blt_strategy_basic(blt_ctx *ctx, blt_strategy **out);
/* PassthroughStrategy (tokenizer.rs:133-145). */
BLT_API int blt_strategy_passthrough(blt_ctx *ctx, blt_strategy **out);
/* BpeStrategy::new over load_bpe_merges_from_path (tokenizer.rs:44-51, config_loader.rs:14-46).
 * Errors carry the reference's kinds and message texts, wrapped like lib.rs:194-201
 * ("Failed to load BPE merges: ...", BLT_ERR_INVALID_INPUT). */
BLT_API int blt_strategy_bpe_from_file(blt_ctx *ctx, const char *merges_path, blt_strategy **out);
/* BpeStrategy::new(Arc<BpeMerges>) for an arbitrary HashMap<(u16,u16),u16> (lib.rs:75,
 * tokenizer.rs:48): n entries (left[i], right[i]) -> value[i]; a later duplicate key overwrites. */
BLT_API int blt_strategy_bpe_from_pairs(blt_ctx *ctx, const uint16_t *left, const uint16_t *right,
                                        const uint16_t *value, size_t n, blt_strategy **out);
BLT_API void blt_strategy_destroy(blt_strategy *s);
/* Number of entries in the strategy's merge map (0 for basic/passthrough). */
BLT_API size_t blt_strategy_num_merges(const blt_strategy *s);

/* ---- the hot path ------------------------------------------------------------------------------ */

/* TokenizationStrategy::process_chunk(&self, &[u8]) -> io::Result<Vec<u8>> (tokenizer.rs:30;
 * called from pipeline.rs:144 and :343).  HOST buffers: `in` is borrowed for the call, the caller
 * owns `out`; out_cap >= 2*n always suffices.  n == 0 -> *out_len = 0 without a launch
 * (tokenizer.rs:57-59).  Re-entrant on one strategy.  Buffers may be ordinary pageable memory (they are
 * staged through pinned buffers inside) or page-locked (cudaHostAlloc / cudaHostRegister: copied directly). */
BLT_API int blt_process_chunk(blt_strategy *s, const uint8_t *in, size_t n, uint8_t *out,
                              size_t out_cap, size_t *out_len);

/* The mmap pipeline of run_mmap_pipeline (pipeline.rs:56-131) on HOST memory: chunk k is bytes
 * [k*chunk, min((k+1)*chunk, n)) (pipeline.rs:73-81); chunks stream through pinned double buffers
 * (H2D, kernel, D2H on separate streams); outputs are concatenated in chunk order
 * (pipeline.rs:153-168).  content_type != NONE writes the 2-byte prefix first (lib.rs:284-294).
 * out_cap >= 2*n + 2 always suffices.  This is the call bench.py times as `e2e`. */
BLT_API int blt_tokenize_host(blt_strategy *s, const uint8_t *in, size_t n, size_t chunk_size,
                              int content_type, uint8_t *out, size_t out_cap, size_t *out_len);

/* Device-resident form of the same pipeline: d_in (16-byte aligned device pointer, n bytes)
 * -> d_out (16-byte aligned, capacity out_cap bytes; 2*n always suffices).  Chunk walls are at
 * multiples of chunk_size exactly as pipeline.rs:73-81 cuts them (chunk_size == 0 or >= n: one
 * chunk).  Work is enqueued on `stream` (a cudaStream_t; NULL = the legacy default stream).
 *   d_chunk_ends : optional device array of ceil(n/chunk_size) uint64; entry k receives the number
 *                  of OUTPUT BYTES produced by chunks 0..k (inclusive prefix), so the last entry is
 *                  the total.  May be NULL.
 *   out_len      : optional HOST pointer; if non-NULL the call synchronises `stream` and stores the
 *                  total output length in bytes.  If NULL the call returns as soon as the work is
 *                  enqueued (use blt_resident_result afterwards).
 * One strategy owns one device workspace: asynchronous calls of one thread on one stream simply queue
 * up behind each other; a call from another thread or on another stream first waits for the
 * outstanding one (its result is kept for the thread that made it).  For truly concurrent device-
 * resident work create one strategy per stream (the table is 128 KiB).
 * Basic mode stores 32 bytes at a time when d_out is 32-byte aligned and 16 at a time otherwise.
 * This is the call bench.py times as `value` (inputs already in HBM). */
BLT_API int blt_process_resident(blt_strategy *s, const void *d_in, size_t n, size_t chunk_size,
                                 void *d_out, size_t out_cap, uint64_t *d_chunk_ends, void *stream,
                                 size_t *out_len);
/* Synchronises `stream` and returns the output length (bytes) and the number of BPE sweeps of the
 * most recent blt_process_resident / blt_detokenize_resident call made on this strategy by the
 * calling thread (also when a later call of another thread had to wait for it in the meantime). */
BLT_API int blt_resident_result(blt_strategy *s, void *stream, size_t *out_len, uint32_t *sweeps);

/* ---- detokenizer: big-endian u16 tokens -> bytes ------------------------------------
 * The reference has no detokenizer; this is the consumer of its wire format (tokenizer.rs:88-91, optional
 * content-type token first, lib.rs:284-294) and the round-trip check of the tokenizer.  Defined for basic,
 * passthrough (a copy) and for BPE tables whose keys are byte pairs and whose ids are >= 256 and distinct
 * (every merges.txt table, config_loader.rs:27-40): token < 256 -> that byte, id of (l, r) -> bytes l r.
 * Errors: BLT_ERR_INVALID_INPUT if the table is not invertible; BLT_ERR_INVALID_DATA for an odd byte
 * count, a missing content-type token or a token that is not in the table; BLT_ERR_CAPACITY.  The output
 * is never longer than the input (out_cap >= n_bytes always suffices). */
BLT_API int blt_detokenize_host(blt_strategy *s, const uint8_t *in, size_t n_bytes, int has_content_type,
                                uint8_t *out, size_t out_cap, size_t *out_len);
/* Device-resident form: d_tokens / d_out are 16-byte aligned device pointers, no content-type token.
 * out_len == NULL: returns when the work is enqueued (blt_resident_result gives the length later). */
BLT_API int blt_detokenize_resident(blt_strategy *s, const void *d_tokens, size_t n_bytes, void *d_out,
                                    size_t out_cap, void *stream, size_t *out_len);

/* ---- merges "training": pair histogram -> merges.txt lines ---------------------------
 * The reference only consumes merges files (config_loader.rs:14-46); its benchmark tables are "the k most
 * frequent adjacent byte pairs" of a sample, which is this histogram: counts[b0 << 8 | b1] = number of i with
 * in[i] = b0 and in[i+1] = b1 (65 536 entries). */
BLT_API int blt_count_pairs_host(blt_ctx *ctx, const uint8_t *in, size_t n, uint64_t *counts);
/* d_in: device pointer, 16-byte aligned; d_counts: device array of 65 536 uint64, overwritten. */
BLT_API int blt_count_pairs_resident(blt_ctx *ctx, const void *d_in, size_t n, uint64_t *d_counts, void *stream);
/* Host: the k most frequent pairs, most frequent first, ties by b0*256+b1 ascending; if fewer than k pairs
 * occur and pad_unobserved != 0, never-observed pairs follow in ascending b0*256+b1.
 * left/right receive *n_out <= k entries: line i of the merges.txt is "left[i] right[i]" (id 256 + i). */
BLT_API int blt_select_merges(const uint64_t *counts, size_t k, int pad_unobserved, uint8_t *left, uint8_t *right,
                              size_t *n_out);

/* ---- run_tokenizer: file to file ---------------------------------------------------------------- */

/* CoreConfig (blt_core/src/lib.rs:110-130) as built by CoreConfig::new_from_cli (lib.rs:149-174),
 * plus the device list.  NULL paths mean stdin / stdout (io_handler.rs:58-61, 74). */
typedef struct blt_core_config {
    const char *input;        /* Option<PathBuf>; NULL = stdin                                       */
    const char *output;       /* Option<PathBuf>; NULL = stdout                                      */
    const char *merges_file;  /* Option<PathBuf>; loaded eagerly, before any IO (lib.rs:160)          */
    int content_type;         /* blt_content_type                                                    */
    int has_threads;          /* Option<usize> threads: 0 = None (all logical CPUs, utils.rs:88-95)   */
    size_t threads;           /*   Some(0) behaves as 1 (utils.rs:81-86)                              */
    const char *chunk_size;   /* Option<String>, e.g. "16MB" (utils.rs:10-45); NULL = auto           */
    int has_memcap;           /* Option<u8>: 0 = None -> 80 (lib.rs:170)                              */
    unsigned memcap;          /*   percent of RAM for the automatic chunk size (chunking.rs:41-42)    */
    int passthrough;          /* bool (lib.rs:129)                                                    */
    int num_gpus;             /* NEW: GPUs to shard chunks over; 0 = one (each costs ~1 s of start-up) */
} blt_core_config;

/* run_tokenizer(CoreConfig::new_from_cli(...)) (lib.rs:149-174, 245-267): parse + load merges, pick
 * the strategy, size chunks, open input (mmap) / output, write the content-type prefix, run the
 * chunk pipeline, write results in chunk order.  With num_gpus = G > 1 the chunks are dealt ROUND-ROBIN: chunk k is
 * processed by GPU blt_file_chunk_device(k, G) = k mod G, each GPU with its own pipeline; the G pipelines advance
 * through the file together, so chunk k's file offset (the sum of the output lengths of chunks 0..k-1, kept on a
 * host-side board) is known almost as soon as its own bytes are back.  Nothing is exchanged between devices. */
BLT_API int blt_run_tokenizer(const blt_core_config *cfg);
/* The inverse, file to file (no reference counterpart): `input` holds big-endian u16 tokens, `output`
 * receives the bytes; merges_file / passthrough select the table as above; content_type != NONE means the
 * stream starts with a content-type token, which is checked and dropped.  One GPU, whole file in memory. */
BLT_API int blt_run_detokenizer(const blt_core_config *cfg);

/* ---- host-only helpers (no device needed) --------------------------------------------------------- */

/* blt_core::load_bpe_merges (lib.rs:216-230) / load_bpe_merges_from_path (config_loader.rs:14-46).
 * Writes up to cap entries sorted by (left,right); *n receives the map size.  Errors are the bare
 * config_loader kinds (NotFound / InvalidData), not the lib.rs:194-201 wrapping. */
BLT_API int blt_load_bpe_merges(const char *path, uint16_t *left, uint16_t *right, uint16_t *value,
                                size_t cap, size_t *n);
/* parse_chunk_size_str (utils.rs:10-45); failure -> BLT_ERR_INVALID_INPUT (lib.rs:176-182). */
BLT_API int blt_parse_chunk_size(const char *s, size_t *out);
/* get_effective_chunk_size (chunking.rs:26-62).  total_ram_bytes == 0 probes the host the way
 * sysinfo::System::total_memory() does (MemTotal). */
BLT_API size_t blt_effective_chunk_size(int has_cli, size_t cli_size, size_t threads, unsigned memcap,
                                        uint64_t total_ram_bytes);
/* determine_thread_count (utils.rs:79-97). */
BLT_API size_t blt_determine_thread_count(int has_override, size_t override_val);
/* ContentType::get_token_value (lib.rs:96-103); 0 for BLT_CONTENT_NONE. */
BLT_API uint16_t blt_content_type_token(int content_type);
/* Contiguous partition for callers that give every GPU (or rank) its OWN output, e.g. one process per GPU:
 * chunk k of K goes to GPU floor(k*G/K); writes the first chunk index of each of the G+1
 * range bounds into bounds[0..G].  blt_run_tokenizer, whose GPUs share ONE ordered output file, does not use it
 * (see blt_file_chunk_device). */
BLT_API void blt_shard_chunks(size_t n_chunks, int n_gpus, size_t *bounds);
/* The device that blt_run_tokenizer gives chunk k when it runs on n_gpus devices: k mod n_gpus. */
BLT_API int blt_file_chunk_device(size_t chunk_index, int n_gpus);

#ifdef __cplusplus
}
#endif
#endif /* BLT_CUDA_H */
