// kernels.cuh -- launch interface of the sm_100a kernels behind libblt_cuda.so.
//
// K1  widen_kernel        BasicTokenizationStrategy::process_chunk  (blt_core/src/tokenizer.rs:108-123)
// K2  dense_pairs_kernel  one sweep of BpeStrategy::process_chunk   (blt_core/src/tokenizer.rs:63-86)
//     + count/scan/emit   in its closed parallel form: a speculative streaming pass for
//                         merge-dense input, and the exact three-launch sweep (pair lookup -> run parity ->
//                         per-range carry functions -> scan -> compaction -> big-endian u16 store).
// K3  the same exact sweep over u16 tokens with a general HashMap<(u16,u16),u16> (lib.rs:75).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace bltk {

// Byte-pair table for K2: 65 536 u16 entries, direct-indexed by a swizzled (b0,b1) key.
// entry = bswap16(id) when (b0,b1) is a rule (requires id >= 256, so the LOW byte of the stored
// value is non-zero), else bswap16(b0) = b0 << 8 (low byte zero).  One shared-memory read therefore
// yields membership AND the big-endian token to emit at that position.
constexpr int kPairTableEntries = 65536;
__host__ __device__ inline uint32_t pair_table_index(uint32_t b0, uint32_t b1) {
    // Bank (= bits 1..5 of the u16 index) mixes both bytes so lanes holding the same frequent
    // first byte with different second bytes do not collide.
    const uint32_t x = b0 | (b1 << 8);
    return x ^ (x >> 7);  // xorshift: a bijection on 16 bits
}

// General map for K3: open addressing, linear probing, 8-byte slots.
struct HashSlot {
    uint32_t key;    // (left << 16) | right
    uint16_t value;  // little-endian token id
    uint16_t used;   // 1 if occupied
};
__host__ __device__ inline uint32_t hash_pair(uint32_t key) {
    key ^= key >> 15; key *= 0x2c1b3c6du; key ^= key >> 12; key *= 0x297a2d39u; key ^= key >> 15;
    return key;
}
// One-hash Bloom filter over the keys (65 536 bits): a pair whose bit is clear is not a rule, so the table in global
// memory is only probed for rules and for a share of n_rules / 65 536 of the other candidate pairs.
__host__ __device__ inline uint32_t pair_bloom_bit(uint32_t left, uint32_t right) {
    const uint32_t h = left * 40503u + right * 60493u;
    return (h ^ (h >> 15)) & 0xFFFFu;
}
struct HashTableView {
    const HashSlot *slots;     // device
    uint32_t mask;             // capacity - 1 (capacity is a power of two, load <= 0.5)
    const uint32_t *can_left;  // device bitmap, 65 536 bits: token appears as a left component
    const uint32_t *can_right; // device bitmap, 65 536 bits: token appears as a right component
    const uint32_t *pair_bloom; // device bitmap, 65 536 bits: pair_bloom_bit(left, right) of every key
    const uint16_t *bytemap;   // device: the rules whose two components are bytes, direct-indexed by pair_table_index:
                               // 65 536 values (host order), then a 65 536-bit "is a rule" bitmap (first sweep: byte input)
};

// Per-launch scratch in device memory (the control block is zeroed by the launcher before every sweep).
struct SweepScratch {
    void *ctrl;              // start of the region: 512-byte control block, then the descriptors
    uint64_t *total_tokens;  // ctrl+0  : number of tokens written by the sweep
    uint32_t *merged_any;    // ctrl+12 : set to 1 if any pair merged in this sweep
    uint32_t *overflow;      // ctrl+16 : set to 1 if the output capacity was exceeded
    uint32_t *dense_abort;   // ctrl+384: set by the dense kernel when its speculation fails (own line)
    uint64_t *tile_status;   // ctrl+512: per warp range: [w] carry_in<<63 | tokens before; [8192+w] tokens of the range
    uint32_t *tile_desc;     // after tile_status: per warp range: carry function flags
    uint8_t *meta;           // after tile_desc: n_elems_max / 16 + 64 bytes (SweepArgs::meta; the fused sweep's tile descriptors)
    size_t meta_bytes;
    size_t bytes;            // size of the whole region
    size_t max_tiles;
};
size_t sweep_scratch_bytes(size_t n_elems_max);
// Carves `mem` (device, >= sweep_scratch_bytes) into a SweepScratch.
SweepScratch sweep_scratch_carve(void *mem, size_t n_elems_max);
// Smallest tile (in input elements) any sweep configuration uses; sizes the status array.
constexpr size_t kMinTileElems = 2048;

struct SweepArgs {
    const void *in;          // device, 16-byte aligned: u8 bytes (K2 / first K3 sweep) or BE u16 tokens
    size_t n;                // number of input elements
    size_t chunk;            // wall every `chunk` elements (0 = none besides the end)
    uint16_t *out;           // device, 16-byte aligned, big-endian u16 tokens
    size_t out_cap_tokens;   // capacity of out in tokens, counted from out[0]
    size_t out_base_tokens;  // the first token of this sweep lands at out[out_base_tokens]
    uint64_t *chunk_ends;    // optional device array: inclusive prefix of OUTPUT BYTES per chunk
    size_t chunk_ends_base;  // bytes added to every chunk_ends entry (output before this launch)
    uint8_t *meta;           // optional (K2 walk variant): one byte per 16-byte segment written by count, read by emit
    SweepScratch scratch;
    size_t total_bias;       // tokens added to the published total (the dense pass's prefix when the exact sweep redoes only the rest)
    uint32_t skip_unmerged_emit;  // K3: a sweep that merges nothing writes nothing (its output would equal its input; the host keeps the input)
};

// Detokenizer (detok.cuh): n_tok big-endian u16 tokens -> bytes.
struct DetokArgs {
    const uint16_t *in;      // device, 16-byte aligned
    size_t n_tok;
    uint8_t *out;            // device, 16-byte aligned
    size_t out_cap;          // bytes
    const uint16_t *table;   // device: 65536 x u16 (id -> l | r << 8; id < 256 -> id), then 2048 x u32 "id exists" bitmap
    uint32_t limit;          // ids >= limit do not exist
    uint32_t holes;          // 1: ids below limit may be missing too (consult the bitmap)
    SweepScratch scratch;    // total_tokens receives the output BYTES; ctrl word 6 = "unknown token seen"
};
// variant 0: three launches (count, scan, emit); 1: the fused single pass (needs n_tok / 2048 + 8 bytes of scratch meta,
// else it falls back to the three launches).  *launches receives the number of kernels launched.
cudaError_t launch_detokenize(const DetokArgs &a, int variant, cudaStream_t stream, int *launches);

// Adjacent-byte-pair histogram (pairhist.cuh): d_counts[b0 << 8 | b1] += occurrences, 65 536 x u64
// (zero_first: cleared by the launch; otherwise accumulated into).
cudaError_t launch_pair_hist(const unsigned char *d_in, size_t n, unsigned long long *d_counts, bool zero_first,
                             cudaStream_t stream);

// K1.  16-byte loads, 32-byte stores.  n bytes in -> 2n bytes out (00 b pairs).
cudaError_t launch_widen(const uint8_t *d_in, size_t n, uint8_t *d_out, cudaStream_t stream);
// chunk_ends[k] = bytes_per_elem * min((k+1)*chunk, n) for the fixed-ratio strategies.
cudaError_t launch_fill_chunk_ends(uint64_t *d_ends, size_t n, size_t chunk, unsigned bytes_per_elem,
                                   cudaStream_t stream);
// K2.  d_table = kPairTableEntries u16 in device memory (layout above).
// try_dense: run the speculative dense pass, which launches the exact sweep itself if it must.
// *host_launches (optional) receives the number of kernels the host enqueued.
cudaError_t launch_bpe_sweep_pairs(const SweepArgs &a, const uint16_t *d_table, int variant, bool try_dense,
                                   cudaStream_t stream, int *host_launches = nullptr);
// K3.  in_is_u16: input is BE u16 tokens (true) or raw bytes (false).
// nb independent sweeps of one input kind in three launches; d_args = device-readable view of h_args (page-locked host memory)
cudaError_t launch_bpe_sweep_hash_batch(const SweepArgs *h_args, const SweepArgs *d_args, int nb, const HashTableView &t, bool in_is_u16,
                                        cudaStream_t stream);
cudaError_t launch_bpe_sweep_hash(const SweepArgs &a, const HashTableView &t, bool in_is_u16,
                                  cudaStream_t stream);
// Kernels the HOST enqueues per K2 call: the dense pass alone when it is attempted (it launches count, scan
// and emit itself, from the device, only if its speculation fails); count + scan + emit otherwise.
constexpr int kLaunchesDenseAttempt = 1, kLaunchesExact = 3, kLaunchesFused = 2;

int num_sweep_variants();
const char *sweep_variant_name(int variant);

}  // namespace bltk
