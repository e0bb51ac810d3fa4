/*
 * blt_oracle.h -- CPU ORACLE for the blt tokenization hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a C++17 restatement of the reference's CPU algorithm (jtrefon/blt v0.2.2, pure Rust).
 * It exists so the CUDA path can be checked bit-for-bit and so a CPU baseline can be timed on the
 * GPU box's host cores.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.  The product (blt_b200/, libblt_cuda.so) never links,
 * imports or calls anything in this directory and has no CPU fallback.
 *
 * PARITY PINNING: the Rust reference cannot be built in this image (no rustc/cargo), so the oracle
 * is pinned on every golden vector the reference's own tests hold for this path
 * (blt_core/src/tokenizer.rs:171-291, blt_core/src/config_loader.rs:62-202,
 * blt_core/src/chunking.rs:98-112, blt_core/src/utils.rs:52-70, tests/cli.rs:20-214); see
 * tests/test_oracle_golden.py.  Behaviour the reference tests do not pin (chunk boundaries, runs,
 * big tables) rests on this restatement plus an independent second model (oracle/py_model.py).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 */
#ifndef BLT_ORACLE_H
#define BLT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors of the std::io::ErrorKind values the reference uses on this path. */
enum {
    ORA_OK = 0,
    ORA_NOT_FOUND = -1,     /* io::ErrorKind::NotFound      (File::open, config_loader.rs:15)          */
    ORA_INVALID_INPUT = -2, /* io::ErrorKind::InvalidInput  (lib.rs:176-182, lib.rs:194-201)           */
    ORA_INVALID_DATA = -3,  /* io::ErrorKind::InvalidData   (config_loader.rs:27-43)                   */
    ORA_IO = -4,            /* any other io::Error                                                     */
    ORA_CAPACITY = -7       /* caller's output buffer too small (no reference counterpart)             */
};

enum { ORA_MODE_BASIC = 0, ORA_MODE_BPE = 1, ORA_MODE_PASSTHROUGH = 2 };

/* BpeMerges = HashMap<(u16,u16),u16>  (blt_core/src/lib.rs:75) */
typedef struct ora_merges ora_merges;

ora_merges *ora_merges_new(void);
void ora_merges_free(ora_merges *m);
/* HashMap::insert semantics: a later insert of the same key overwrites (config_loader.rs:39). */
void ora_merges_insert(ora_merges *m, uint16_t left, uint16_t right, uint16_t value);
size_t ora_merges_len(const ora_merges *m);
/* Export entries sorted by (left,right); returns the number written (<= cap). */
size_t ora_merges_export(const ora_merges *m, uint16_t *left, uint16_t *right, uint16_t *value,
                         size_t cap);

/* load_bpe_merges_from_path (blt_core/src/config_loader.rs:14-46).  On error *out is NULL and
 * err holds the io::Error message text. */
int ora_load_bpe_merges(const char *path, ora_merges **out, char *err, size_t errcap);

/* BpeStrategy::process_chunk (blt_core/src/tokenizer.rs:56-93).  *sweeps (optional) receives the
 * number of passes of the outer loop, including the final pass that merges nothing. */
int ora_bpe_process_chunk(const ora_merges *m, const uint8_t *in, size_t n, uint8_t *out,
                          size_t out_cap, size_t *out_len, uint32_t *sweeps);
/* BasicTokenizationStrategy::process_chunk (blt_core/src/tokenizer.rs:108-123). */
int ora_basic_process_chunk(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                            size_t *out_len);
/* PassthroughStrategy::process_chunk (blt_core/src/tokenizer.rs:138-144). */
int ora_passthrough_process_chunk(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                                  size_t *out_len);

/* run_tokenizer on an in-memory "mmap": content-type prefix (lib.rs:284-294), fixed-offset chunk
 * table (pipeline.rs:73-81), <= threads chunks in flight (pipeline.rs:85-101), results written in
 * chunk order (pipeline.rs:153-168).  content_type_token < 0 means none. */
int ora_run_buffer(int mode, const ora_merges *m, const uint8_t *in, size_t n, size_t chunk_size,
                   size_t threads, int content_type_token, uint8_t *out, size_t out_cap,
                   size_t *out_len);
/* Same, file to file: input mmap'ed read-only (io_handler.rs:53-57), output created/truncated and
 * written through a buffered writer (io_handler.rs:68-72). */
int ora_run_files(int mode, const ora_merges *m, const char *in_path, const char *out_path,
                  size_t chunk_size, size_t threads, int content_type_token, char *err,
                  size_t errcap);

/* parse_chunk_size_str (blt_core/src/utils.rs:10-45). */
int ora_parse_chunk_size(const char *s, size_t *out, char *err, size_t errcap);
/* get_effective_chunk_size (blt_core/src/chunking.rs:26-62); total_ram_bytes stands in for
 * sysinfo's total_memory(). */
size_t ora_effective_chunk_size(int has_cli, size_t cli_size, size_t threads, unsigned memcap,
                                uint64_t total_ram_bytes);
/* determine_thread_count (blt_core/src/utils.rs:79-97); logical_cpus stands in for num_cpus. */
size_t ora_determine_thread_count(int has_override, size_t override_val, size_t logical_cpus);
/* ContentType::get_token_value (blt_core/src/lib.rs:96-103): 0 text,1 audio,2 bin,3 video. */
uint16_t ora_content_type_token(int content_type);

/* Detokenizer: the inverse of the wire format (big-endian u16 tokens, tokenizer.rs:88-91; optional
 * content-type token first, lib.rs:284-294).  THE REFERENCE HAS NO DETOKENIZER: this
 * is the checker of the GPU detokenizer and of round trips, defined only for tables whose keys are byte
 * pairs and whose ids are >= 256 and distinct (every merges.txt table, config_loader.rs:27-40).
 * token < 256 -> that byte; token = id of (l, r) -> bytes l r.  Errors: ORA_INVALID_INPUT if the table
 * is not invertible, ORA_INVALID_DATA for an odd byte count, a missing/unknown content-type token or a
 * token that is not in the table. */
int ora_detokenize(const ora_merges *m, const uint8_t *in, size_t n_bytes, int has_content_type,
                   uint8_t *out, size_t out_cap, size_t *out_len);

#ifdef __cplusplus
}
#endif
#endif
