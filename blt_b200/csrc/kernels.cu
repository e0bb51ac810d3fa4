// kernels.cu -- hand-written sm_100a kernels for blt's tokenization hot path.  See DESIGN.md.
//
// The sweep kernel implements ONE pass of the loop at blt_core/src/tokenizer.rs:63-86 in its closed
// parallel form.  With t[] the input tokens and
//     m[i]     = 1 iff (t[i], t[i+1]) is a key of the merge map, 0 at every chunk-last index,
//     start[i] = m[i] & ~start[i-1]                       (start[-1] = 0 at every chunk start)
// the reference's greedy left-to-right scan emits map[(t[i],t[i+1])] where start[i], drops token i
// where start[i-1], and copies it otherwise.  Inside a maximal run of m = 1 the starts are the
// positions at even distance from the run's first position, so a 16-element segment acts on the
// incoming carry (= "my first element was consumed by the previous segment") either as the
// identity (m all ones) or as a constant (m has a zero).
//
// Files: this one holds the streaming kernels (widen_kernel, dense_pairs_kernel), the two lookup front ends
// and the launchers; sweep3.cuh the exact sweep (count / scan / emit, no inter-CTA waiting); detok.cuh the
// detokenizer; pairhist.cuh the pair histogram.  Everything is compiled as relocatable device code because
// the dense pass launches the exact sweep from the device when its speculation fails.
#include "kernels.cuh"

#include <algorithm>
#include <atomic>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace bltk {
namespace {

constexpr int kCtaThreads = 1024;
constexpr uint32_t FULL = 0xffffffffu;
constexpr int kMaxDevices = 64;
constexpr size_t kCtrlBytes = 512;  // results | (spare) | the dense pass's work counter | its abort word, 128 bytes each
constexpr uint32_t kDenseUnitSegs = 512;  // 8 KiB of input per warp and work unit

int sm_count(int dev) {
    static std::atomic<int> cached[kMaxDevices];
    int v = cached[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cached[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// ---- small PTX helpers -------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_v4(void *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
// 32-byte store (sm_100+, PTX 8.8): one instruction per 16 input bytes in the widen kernel.
__device__ __forceinline__ void stg_v8(void *p, const uint4 &a, const uint4 &b) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y),
                 "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

// ================================================================================================
// K1: byte -> big-endian u16 (tokenizer.rs:108-123): output bytes are 00 b.
// ================================================================================================
__device__ __forceinline__ uint4 widen8(uint32_t lo, uint32_t hi) {
    // little-endian words whose memory image is 00 b0 00 b1 | 00 b2 00 b3 | ...
    uint4 r;
    r.x = __byte_perm(lo, 0, 0x1404);
    r.y = __byte_perm(lo, 0, 0x3424);
    r.z = __byte_perm(hi, 0, 0x1404);
    r.w = __byte_perm(hi, 0, 0x3424);
    return r;
}

// V8: one 32-byte store per 16 input bytes (needs a 32-byte aligned output); otherwise two 16-byte stores (the
// C ABI only promises 16-byte alignment, e.g. a slice of a larger device buffer).
template <bool V8>
__device__ __forceinline__ void widen_store(uint8_t *p, const uint4 &a) {
    if (V8) {
        stg_v8(p, widen8(a.x, a.y), widen8(a.z, a.w));
    } else {
        stg_stream_v4(p, widen8(a.x, a.y));
        stg_stream_v4(p + 16, widen8(a.z, a.w));
    }
}

template <bool V8>
__global__ void __launch_bounds__(256) widen_kernel(const uint8_t *__restrict__ in, size_t n,
                                                    uint8_t *__restrict__ out) {
    const size_t nvec = n / 16;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    // two independent 16-byte loads in flight per thread per trip
    for (; i + stride < nvec; i += 2 * stride) {
        const uint4 a = ldg_stream_v4(in + i * 16);
        const uint4 b = ldg_stream_v4(in + (i + stride) * 16);
        widen_store<V8>(out + i * 32, a);
        widen_store<V8>(out + (i + stride) * 32, b);
    }
    if (i < nvec) {
        const uint4 a = ldg_stream_v4(in + i * 16);
        widen_store<V8>(out + i * 32, a);
    }
    // ragged tail (< 16 bytes), one thread
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (size_t k = nvec * 16; k < n; ++k) {
            out[2 * k] = 0;
            out[2 * k + 1] = in[k];
        }
    }
}

// chunk_ends for the fixed-ratio strategies (basic: 2 bytes per input byte, passthrough: 1).
__global__ void fill_chunk_ends_kernel(uint64_t *ends, size_t n_chunks, size_t n, size_t chunk, unsigned bytes_per_elem) {
    const size_t k = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < n_chunks) {
        const size_t end = (k + 1 == n_chunks) ? n : (k + 1) * chunk;
        ends[k] = uint64_t(end) * bytes_per_elem;
    }
}

// start bits of one segment: m = pair-membership bits, cin = first element already consumed
__device__ __forceinline__ uint32_t start_bits(uint32_t m, uint32_t cin) {
    const uint32_t mm = m & ~cin;
    const uint32_t s = mm & ~(mm << 1);                      // first bit of every run of ones
    const uint32_t e = mm & ~(mm + (s & 0x55555555u));       // runs that begin at an even position
    return (e & 0x55555555u) | (mm & ~e & 0xAAAAAAAAu);      // same parity as the run's first bit
}

// spread the low 8 (or 4) bits of x to the even bit positions
__device__ __forceinline__ uint32_t spread_even(uint32_t x) {
    x = (x | (x << 4)) & 0x0F0Fu;
    x = (x | (x << 2)) & 0x3333u;
    x = (x | (x << 1)) & 0x5555u;
    return x;
}

// ================================================================================================
// Front ends.  lookup_half() looks up the SEG/2 pairs that START at positions of parity `par` of one
// 16-byte segment and returns (a) in vals[] the big-endian u16 to emit at each of those positions
// (merged id if the pair is a rule, else the element itself), two per register in position order,
// and (b) a SEG/2-bit membership mask (bit i <-> position 2i+par).  all_present() answers "were all
// of them rules" without the mask when the front end can tell from the values alone.
// ================================================================================================

// K2 front end: byte input, direct-indexed byte-pair table in shared memory.
struct PairsFE {
    static constexpr int SEG = 16;       // elements per 16-byte segment
    static constexpr int ELEM = 1;       // bytes per element
    static constexpr int TABLE_BYTES = kPairTableEntries * 2;
    static constexpr bool kMembershipInValue = true;  // present <=> low byte of the stored value != 0
    struct Params { const uint16_t *table; };
    const unsigned char *tbl;            // shared memory

    __device__ __forceinline__ void init(const Params &p, unsigned char *smem) {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.table);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < TABLE_BYTES / 16; i += blockDim.x) dst[i] = src[i];
        tbl = smem;
    }
    __device__ __forceinline__ static uint32_t first_elem(const uint4 &w) { return w.x & 0xffu; }
    __device__ __forceinline__ static uint32_t raw_be(const uint4 &w, int j) {
        const uint32_t word = (j < 4) ? w.x : (j < 8) ? w.y : (j < 12) ? w.z : w.w;
        return ((word >> (8 * (j & 3))) & 0xffu) << 8;  // bswap16(byte)
    }
    __device__ __forceinline__ static uint32_t load_elem(const void *in, size_t pos) {
        return static_cast<const uint8_t *>(in)[pos];
    }
    // Values only.  Two lookups share the xorshift: Y = W ^ ((W >> 7) & 0x01FF01FF) applies
    // idx = x ^ (x >> 7) (== pair_table_index) to both 16-bit halves of the word at once.
    __device__ __forceinline__ void lookup_vals(const uint4 &w, uint32_t next, uint32_t par, uint32_t *vals) const {
        const uint32_t sh = par * 8;  // shift the byte window by one for the odd parity
        const uint32_t W[4] = {__funnelshift_r(w.x, w.y, sh), __funnelshift_r(w.y, w.z, sh),
                               __funnelshift_r(w.z, w.w, sh), __funnelshift_r(w.w, next, sh)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t y = W[k] ^ ((W[k] >> 7) & 0x01FF01FFu);
            const uint32_t e0 = *reinterpret_cast<const uint16_t *>(tbl + ((y << 1) & 0x1FFFEu));
            const uint32_t e1 = *reinterpret_cast<const uint16_t *>(tbl + ((y >> 15) & 0x1FFFEu));
            vals[k] = e0 | (e1 << 16);
        }
    }
    __device__ __forceinline__ static bool all_present(const uint32_t *vals) {
        // (b + 0xFF) carries into bit 8 iff the low byte b of a half is non-zero
        const uint32_t c = 0x00FF00FFu;
        const uint32_t t = ((vals[0] & c) + c) & ((vals[1] & c) + c) & ((vals[2] & c) + c) & ((vals[3] & c) + c);
        return (t & 0x01000100u) == 0x01000100u;
    }
    // bit i <-> token i present.  The low bytes of four tokens are gathered into one word (PRMT), a SWAR
    // "byte != 0" sets bit 7 of each byte, and one multiply collects those four bits.
    __device__ __forceinline__ static uint32_t nonzero_bytes4(uint32_t x) {
        const uint32_t y = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;
        return (((y >> 7) & 0x01010101u) * 0x01020408u) >> 24;
    }
    __device__ __forceinline__ static uint32_t present_mask(const uint32_t *vals) {
        return nonzero_bytes4(__byte_perm(vals[0], vals[1], 0x6420)) | (nonzero_bytes4(__byte_perm(vals[2], vals[3], 0x6420)) << 4);
    }
    __device__ __forceinline__ uint32_t lookup_half(const uint4 &w, uint32_t next, uint32_t par, uint32_t *vals) const {
        lookup_vals(w, next, par, vals);
        return present_mask(vals);
    }
};

// K3 front end: general HashMap<(u16,u16),u16> in global memory (L2-resident), prefiltered by three
// 8 KiB shared-memory bitmaps ("can be a left / right component", and a Bloom filter over the pairs, without which
// the misses among the candidate pairs - most of them on text - each cost dependent global loads).
// Input is raw bytes (first sweep) or big-endian u16 tokens.
template <bool IN_U16>
struct HashFE {
    static constexpr int SEG = IN_U16 ? 8 : 16;
    static constexpr int ELEM = IN_U16 ? 2 : 1;
    static constexpr int TABLE_BYTES = 3 * 8192;
    static constexpr bool kMembershipInValue = false;  // ids may be < 256 and may equal the element
    struct Params { HashTableView t; };
    const uint32_t *can_left, *can_right, *bloom;  // shared memory
    const HashSlot *slots;
    uint32_t mask;

    __device__ __forceinline__ void init(const Params &p, unsigned char *smem) {
        uint32_t *s = reinterpret_cast<uint32_t *>(smem);
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
            s[i] = p.t.can_left[i];
            s[2048 + i] = p.t.can_right[i];
            s[4096 + i] = p.t.pair_bloom[i];
        }
        can_left = s;
        can_right = s + 2048;
        bloom = s + 4096;
        slots = p.t.slots;
        mask = p.t.mask;
    }
    // element j of the segment in host order; j == SEG is the look-ahead element
    __device__ __forceinline__ static uint32_t elem(const uint4 &w, uint32_t next, int j) {
        if (j >= SEG) return next;
        if (IN_U16) {
            const uint32_t word = (j < 2) ? w.x : (j < 4) ? w.y : (j < 6) ? w.z : w.w;
            return __byte_perm((word >> (16 * (j & 1))) & 0xffffu, 0, 0x4401);  // bswap16
        }
        const uint32_t word = (j < 4) ? w.x : (j < 8) ? w.y : (j < 12) ? w.z : w.w;
        return (word >> (8 * (j & 3))) & 0xffu;
    }
    __device__ __forceinline__ static uint32_t first_elem(const uint4 &w) { return elem(w, 0, 0); }
    __device__ __forceinline__ static uint32_t raw_be(const uint4 &w, int j) {
        return __byte_perm(elem(w, 0, j), 0, 0x4401);
    }
    __device__ __forceinline__ static uint32_t load_elem(const void *in, size_t pos) {
        if (IN_U16) return __byte_perm(uint32_t(static_cast<const uint16_t *>(in)[pos]), 0, 0x4401);
        return static_cast<const uint8_t *>(in)[pos];
    }
    __device__ __forceinline__ uint32_t lookup_one(uint32_t cur, uint32_t nxt, uint32_t *out) const {
        *out = cur;
        const uint32_t bb = pair_bloom_bit(cur, nxt);
        if (((can_left[cur >> 5] >> (cur & 31)) & (can_right[nxt >> 5] >> (nxt & 31)) & (bloom[bb >> 5] >> (bb & 31)) & 1u)) {
            const uint32_t key = (cur << 16) | nxt;
            uint32_t h = hash_pair(key) & mask;
            for (;;) {
                const HashSlot s = slots[h];
                if (!s.used) break;
                if (s.key == key) { *out = s.value; return 1u; }
                h = (h + 1) & mask;
            }
        }
        return 0u;
    }
    __device__ __forceinline__ uint32_t lookup_half(const uint4 &w, uint32_t next, uint32_t par, uint32_t *vals) const {
        uint32_t present = 0;
#pragma unroll
        for (int i = 0; i < SEG / 2; ++i) {
            const uint32_t cur = par ? elem(w, next, 2 * i + 1) : elem(w, next, 2 * i);
            const uint32_t nxt = par ? elem(w, next, 2 * i + 2) : elem(w, next, 2 * i + 1);
            uint32_t out;
            present |= lookup_one(cur, nxt, &out) << i;
            const uint32_t be = __byte_perm(out, 0, 0x4401);
            if (i & 1) vals[i >> 1] |= be << 16; else vals[i >> 1] = be;
        }
        return present;
    }
    __device__ __forceinline__ void lookup_vals(const uint4 &, uint32_t, uint32_t, uint32_t *) const {}
    __device__ __forceinline__ static bool all_present(const uint32_t *) { return false; }
    __device__ __forceinline__ static uint32_t present_mask(const uint32_t *) { return 0; }
};

// K3 front end of the FIRST sweep (byte input): only rules whose two components are bytes can match there, and those
// fit a direct table in shared memory - 65 536 values plus an exact "is a rule" bitmap (a value says nothing about
// membership here: ids may be < 256 and may equal the element).  No probe of the hash table, no divergence.
struct ByteMapFE {
    using H = HashFE<false>;
    static constexpr int SEG = 16;
    static constexpr int ELEM = 1;
    static constexpr int TABLE_BYTES = kPairTableEntries * 2 + 8192;
    static constexpr bool kMembershipInValue = false;
    struct Params { const uint16_t *bytemap; };
    const uint16_t *vals16;   // shared memory
    const uint32_t *member;

    __device__ __forceinline__ void init(const Params &p, unsigned char *smem) {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.bytemap);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < TABLE_BYTES / 16; i += blockDim.x) dst[i] = src[i];
        vals16 = reinterpret_cast<const uint16_t *>(smem);
        member = reinterpret_cast<const uint32_t *>(smem + kPairTableEntries * 2);
    }
    __device__ __forceinline__ static uint32_t first_elem(const uint4 &w) { return H::first_elem(w); }
    __device__ __forceinline__ static uint32_t raw_be(const uint4 &w, int j) { return H::raw_be(w, j); }
    __device__ __forceinline__ static uint32_t load_elem(const void *in, size_t pos) { return H::load_elem(in, pos); }
    __device__ __forceinline__ uint32_t lookup_half(const uint4 &w, uint32_t next, uint32_t par, uint32_t *vals) const {
        uint32_t present = 0;
#pragma unroll
        for (int i = 0; i < SEG / 2; ++i) {
            const uint32_t cur = par ? H::elem(w, next, 2 * i + 1) : H::elem(w, next, 2 * i);
            const uint32_t nxt = par ? H::elem(w, next, 2 * i + 2) : H::elem(w, next, 2 * i + 1);
            const uint32_t idx = pair_table_index(cur, nxt);
            const uint32_t hit = (member[idx >> 5] >> (idx & 31)) & 1u;
            const uint32_t out = hit ? uint32_t(vals16[idx]) : cur;
            present |= hit << i;
            const uint32_t be = __byte_perm(out, 0, 0x4401);
            if (i & 1) vals[i >> 1] |= be << 16; else vals[i >> 1] = be;
        }
        return present;
    }
    __device__ __forceinline__ void lookup_vals(const uint4 &, uint32_t, uint32_t, uint32_t *) const {}
    __device__ __forceinline__ static bool all_present(const uint32_t *) { return false; }
    __device__ __forceinline__ static uint32_t present_mask(const uint32_t *) { return 0; }
};

// ================================================================================================
// Tile bookkeeping shared by the exact sweep kernels (sweep3.cuh).
// ================================================================================================
struct TileInfo {
    unsigned long long rem0;   // tile_base % chunk (tile_base itself when there are no walls)
    unsigned long long ck0;    // tile_base / chunk
};

template <int SEG>
struct Walls {
    uint32_t endm;            // chunk-last positions inside the segment (incl. the last element n-1)
    unsigned long long ck;    // chunk index of the segment's first element
};

// Which positions of the segment at element offset `off` of the tile are chunk-last.
template <int SEG, int TILE_ELEMS>
__device__ __forceinline__ Walls<SEG> seg_walls(const SweepArgs &a, const TileInfo &ti, uint32_t off,
                                                unsigned long long g) {
    Walls<SEG> w;
    w.endm = 0;
    w.ck = 0;
    if (a.chunk != 0) {
        if (a.chunk >= size_t(TILE_ELEMS)) {  // at most one wall per tile
            unsigned long long rem = ti.rem0 + off;
            w.ck = ti.ck0;
            if (rem >= a.chunk) { rem -= a.chunk; w.ck += 1; }
            const unsigned long long d = a.chunk - 1 - rem;
            if (d < SEG) w.endm = 1u << uint32_t(d);
        } else {  // tiny chunks (tests): walk the segment
            const uint32_t c = uint32_t(a.chunk);
            const uint32_t lin = uint32_t(ti.rem0) + off;
            w.ck = ti.ck0 + lin / c;
            uint32_t r = lin % c;
#pragma unroll
            for (int j = 0; j < SEG; ++j) {
                if (++r == c) { w.endm |= 1u << j; r = 0; }
            }
        }
    }
    if (g < a.n && a.n - 1 - g < SEG) w.endm |= 1u << uint32_t(a.n - 1 - g);  // end of the last chunk
    return w;
}

#include "sweep3.cuh"
#include "fused.cuh"

// ================================================================================================
// K2-dense: the speculative streaming form of the sweep for merge-dense input.
//
// If every pair that starts at an EVEN offset of its chunk is a rule, the reference's scan merges
// exactly those pairs (start[0] = 1, start[1] = 0, start[2] = 1, ...), whatever the odd pairs are: the
// output is the looked-up id of every even pair, token k of the launch sits at out[k], and no carry
// or count has to cross a tile.  With an even chunk size no such pair straddles a wall.  The kernel
// streams under that hypothesis - 16 bytes in, 8 lookups, 16 bytes out per lane, no barrier, no
// look-back - and records the work unit of a pair that is not a rule in the abort word (the smallest such unit
// wins).  Units in front of the failed unit's chunk are still completed, the ones from that chunk on are given up.
// The last CTA to finish then either publishes the totals or launches the exact sweep (sweep3.cuh) from the device
// for the REST of the launch, from the first failed chunk on (the chunks in front of it are dense: their output
// and chunk ends stand, and the sweep's total is biased by their tokens).
// ================================================================================================
__global__ void __launch_bounds__(kCtaThreads, 1)
dense_pairs_kernel(const SweepArgs a, const uint16_t *__restrict__ table, int variant, unsigned exact_grid) {
    extern __shared__ __align__(16) unsigned char smem[];
    PairsFE fe;
    PairsFE::Params p{table};
    fe.init(p, smem);
    __syncthreads();
    const unsigned char *__restrict__ in = static_cast<const unsigned char *>(a.in);
    const unsigned long long n = a.n;
    uint16_t *__restrict__ out = a.out + a.out_base_tokens;
    uint32_t *const work_counter = reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(a.scratch.ctrl) + 256);
    uint32_t *const done_counter = work_counter + 1;
    uint32_t *const abort_flag = a.scratch.dense_abort;
    // Work is handed out per warp in units of kDenseUnitSegs consecutive 16-byte segments: the first unit
    // of a warp is static, the following ones come from one global counter (an SM that gets less memory
    // bandwidth than its neighbours simply takes fewer units; with a static split the slowest SM set the
    // kernel time).  A unit is TRIPS trips of U independent 16-byte loads per lane; the loads of the next
    // trip - and the counter fetch of the next unit - are issued before the current trip is looked up and
    // stored, so a warp always has U..2U loads in flight and never waits for the counter.
    constexpr int U = 4;
    constexpr uint32_t TRIPS = kDenseUnitSegs / (U * 32);
    const unsigned long long n_segs = n / 16;
    const uint32_t full_units = uint32_t(n_segs / kDenseUnitSegs);
    const int lane = threadIdx.x & 31;
    const uint32_t n_static = gridDim.x * (kCtaThreads / 32);
    bool bad = false;
    // The abort word holds ~unit of the smallest failed unit (0: none; atomicMax).  When the chunk size is a whole number
    // of units, a failure only stops the units from the first unit of the failed unit's chunk on (`floor`).
    const unsigned long long ck = (a.chunk == 0 || a.chunk > n) ? n : a.chunk;
    const uint32_t units_per_chunk = (ck % (kDenseUnitSegs * 16ull) == 0 && ck / (kDenseUnitSegs * 16ull) < 0x7fffffffull)
                                         ? uint32_t(ck / (kDenseUnitSegs * 16ull)) : 0u;  // 0: any failure stops everything
    auto floor_of = [&](uint32_t word) -> uint32_t {  // the first unit that is given up
        if (word == 0u) return 0xffffffffu;
        return units_per_chunk ? (~word / units_per_chunk) * units_per_chunk : 0u;
    };
    uint32_t ab = 0;       // the abort word as it was one trip ago: never waited for inside a trip
    uint32_t nxt_raw = 0;  // lane 0: the unit after the current one (fetched one unit ahead)
    if (lane == 0) nxt_raw = atomicAdd(work_counter, 1u) + n_static;
    struct Pos { uint32_t unit, t; };
    auto advance = [&](Pos q) -> Pos {
        if (q.t + 1 < TRIPS) return Pos{q.unit, q.t + 1};
        const uint32_t u = __shfl_sync(FULL, nxt_raw, 0);
        if (lane == 0 && u < full_units) nxt_raw = atomicAdd(work_counter, 1u) + n_static;
        return Pos{u, 0};
    };
    auto seg_of = [&](Pos q) { return (unsigned long long)q.unit * kDenseUnitSegs + q.t * (U * 32) + lane; };
    auto load = [&](uint4 *w, unsigned long long s0) {
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = ldg_stream_v4(in + (s0 + u * 32) * 16);
    };
    // looks up and stores one trip of unit `unit`; true = this warp gives up (the unit failed, or a unit of its chunk or
    // of a chunk in front of it has)
    auto work = [&](const uint4 *w, unsigned long long s0, uint32_t unit) {
        const uint32_t ab_now = ab;
        if (lane == 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(ab) : "l"(abort_flag) : "memory");
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t v[4];
            fe.lookup_vals(w[u], 0u, 0u, v);
            bad = bad || !PairsFE::all_present(v);
            stg_stream_v4(out + (s0 + u * 32) * 8, make_uint4(v[0], v[1], v[2], v[3]));
        }
        if (__any_sync(FULL, bad)) {
            if (lane == 0) atomicMax(abort_flag, ~unit);
            return true;
        }
        return unit >= floor_of(__shfl_sync(FULL, ab_now, 0));
    };
    uint4 wa[U], wb[U];
    Pos pa{blockIdx.x * (kCtaThreads / 32) + (threadIdx.x >> 5), 0}, pb{0, 0};
    bool more = pa.unit < full_units;
    if (more) load(wa, seg_of(pa));
    while (more) {
        pb = advance(pa);
        const bool next = pb.unit < full_units;
        if (next) load(wb, seg_of(pb));
        if (work(wa, seg_of(pa), pa.unit) || !next) break;
        pa = advance(pb);
        more = pa.unit < full_units;
        if (more) load(wa, seg_of(pa));
        if (work(wb, seg_of(pb), pb.unit)) break;
    }
    // the segments behind the last full unit (fewer than kDenseUnitSegs)
    if (blockIdx.x == gridDim.x - 1 && !bad) {
        for (unsigned long long seg = (unsigned long long)full_units * kDenseUnitSegs + threadIdx.x; seg < n_segs;
             seg += kCtaThreads) {
            const uint4 w0 = ldg_stream_v4(in + seg * 16);
            uint32_t v[4];
            fe.lookup_vals(w0, 0u, 0u, v);
            bad = bad || !PairsFE::all_present(v);
            stg_stream_v4(out + seg * 8, make_uint4(v[0], v[1], v[2], v[3]));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // the last n % 16 elements
        unsigned long long i = n_segs * 16;
        for (; i + 1 < n; i += 2) {
            const uint32_t e = *reinterpret_cast<const uint16_t *>(fe.tbl + 2 * pair_table_index(in[i], in[i + 1]));
            if ((e & 0xffu) == 0) bad = true;
            out[i / 2] = uint16_t(e);
        }
        if (i < n) out[i / 2] = uint16_t(uint32_t(in[i]) << 8);  // odd length: the last element stays a raw token
    }
    // (a warp that failed inside a unit has recorded the unit already; what is left is the two tails, which belong to
    // the last chunk: one atomic per CTA - 150 000 threads storing to one word serialise in L2 for tens of microseconds)
    __shared__ uint32_t s_last;
    const bool cta_bad = __syncthreads_or(bad) != 0;
    if (threadIdx.x == 0) {
        if (cta_bad) atomicMax(abort_flag, ~full_units);
        __threadfence();
        s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last == 0) return;
    // ---- the last CTA to finish: everybody else's stores (output and abort word) are visible ----
    __threadfence();
    const uint32_t ab_word = *reinterpret_cast<volatile uint32_t *>(abort_flag);
    const bool failed = ab_word != 0u;
    const unsigned long long c = (a.chunk == 0 || a.chunk > n) ? n : a.chunk;
    const unsigned long long n_chunks = (n + c - 1) / c;
    if (!failed) {  // the speculation held: the output is complete, publish the totals of the sweep
        const unsigned long long tokens = (n + 1) / 2;
        if (threadIdx.x == 0) {
            *a.scratch.total_tokens = tokens;
            *a.scratch.merged_any = (n >= 2) ? 1u : 0u;
            *a.scratch.overflow = 0u;
            reinterpret_cast<uint32_t *>(a.scratch.ctrl)[5] = 0u;  // "dense pass failed" (read by the host's predictor)
        }
        if (a.chunk_ends != nullptr) {
            for (unsigned long long k = threadIdx.x; k < n_chunks; k += blockDim.x)
                a.chunk_ends[k] = a.chunk_ends_base + ((k + 1 == n_chunks) ? 2 * tokens : (k + 1) * c);
        }
    } else {
        // Some even pair is not a rule: the exact sweep redoes the launch from the first failed chunk on.  It is enqueued
        // from here (tail launch: it starts when this grid has completed) so that the host never enqueues - and the GPU
        // never schedules - three kernels that would have nothing to do in the common case.
        // Every unit in front of chunk j was completed without a failure (a unit is only given up when a unit of its own
        // chunk or of a chunk in front of it has failed, and j is the chunk of the smallest failed unit).
        unsigned long long j = units_per_chunk ? (unsigned long long)(~ab_word / units_per_chunk) : 0ull;
        if (j >= n_chunks) j = n_chunks - 1;  // (the tails' pseudo unit)
        if (a.chunk_ends != nullptr) {
            for (unsigned long long k = threadIdx.x; k < j; k += blockDim.x) a.chunk_ends[k] = a.chunk_ends_base + (k + 1) * c;
        }
        if (threadIdx.x < 64) reinterpret_cast<uint32_t *>(a.scratch.ctrl)[threadIdx.x] = 0u;
        __syncthreads();
        if (threadIdx.x == 0) {
            reinterpret_cast<uint32_t *>(a.scratch.ctrl)[5] = 1u;
            reinterpret_cast<uint32_t *>(a.scratch.ctrl)[7] = uint32_t(j * c * 1000ull / n);  // per mille of the launch the dense prefix covers (predictor)
            __threadfence();
            SweepArgs r = a;  // the rest: chunks j .. (c is a multiple of 8 KiB when j != 0: alignment and parity hold)
            r.in = in + j * c;
            r.n = n - j * c;
            r.out_base_tokens = a.out_base_tokens + j * c / 2;
            if (a.chunk_ends != nullptr) r.chunk_ends = a.chunk_ends + j;
            r.chunk_ends_base = a.chunk_ends_base + j * c;
            r.total_bias = a.total_bias + j * c / 2;
            const bool ok = (variant == 1)   ? tail_launch_sweep3<PairsFE, 8>(r, p, exact_grid)
                            : (variant == 2) ? tail_launch_sweep3<PairsFE, 4, true>(r, p, exact_grid)
                                             : tail_launch_sweep3<PairsFE, 4>(r, p, exact_grid);
            if (!ok) *a.scratch.overflow = 2u;  // reported as a CUDA error by the host (never seen so far)
        }
    }
}


#include "detok.cuh"
#include "pairhist.cuh"

}  // namespace

cudaError_t launch_pair_hist(const unsigned char *d_in, size_t n, unsigned long long *d_counts, bool zero_first,
                             cudaStream_t stream) {
    return launch_pair_hist_impl(d_in, n, d_counts, zero_first, stream);
}

cudaError_t launch_detokenize(const DetokArgs &a, int variant, cudaStream_t stream, int *launches) {
    return launch_detok_impl(a, variant, stream, launches);
}

#ifdef BLT_FUSED_PROF
cudaError_t debug_fused_profile(unsigned long long *host_out, size_t n_words) {
    return cudaMemcpyFromSymbol(host_out, g_fz_prof, n_words * 8, 0, cudaMemcpyDeviceToHost);
}
#endif

// ---- scratch -------------------------------------------------------------------------------------
size_t sweep_scratch_bytes(size_t n_elems_max) {
    const size_t tiles = 2 * 8192;
    return kCtrlBytes + tiles * 8 + tiles * 4 + n_elems_max / 16 + 64;
}
SweepScratch sweep_scratch_carve(void *mem, size_t n_elems_max) {
    SweepScratch s;
    const size_t tiles = 2 * 8192;
    unsigned char *p = static_cast<unsigned char *>(mem);
    s.ctrl = p;
    s.total_tokens = reinterpret_cast<uint64_t *>(p);
    s.merged_any = reinterpret_cast<uint32_t *>(p + 12);
    s.overflow = reinterpret_cast<uint32_t *>(p + 16);
    s.dense_abort = reinterpret_cast<uint32_t *>(p + 384);
    s.tile_status = reinterpret_cast<uint64_t *>(p + kCtrlBytes);
    s.tile_desc = reinterpret_cast<uint32_t *>(p + kCtrlBytes + tiles * 8);
    s.meta = p + kCtrlBytes + tiles * 8 + tiles * 4;
    s.meta_bytes = n_elems_max / 16 + 64;
    s.bytes = sweep_scratch_bytes(n_elems_max);
    s.max_tiles = tiles;
    return s;
}

// ---- launchers -------------------------------------------------------------------------------------
cudaError_t launch_widen(const uint8_t *d_in, size_t n, uint8_t *d_out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    const int sms = sm_count(dev % kMaxDevices);
    const size_t nvec = n / 16;
    size_t blocks = (nvec + 255) / 256;
    const size_t cap = size_t(sms) * 8;  // 8 resident CTAs of 256 threads per SM, grid-stride beyond
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    if ((reinterpret_cast<uintptr_t>(d_out) & 31u) == 0) widen_kernel<true><<<dim3(unsigned(blocks)), dim3(256), 0, stream>>>(d_in, n, d_out);
    else widen_kernel<false><<<dim3(unsigned(blocks)), dim3(256), 0, stream>>>(d_in, n, d_out);
    return cudaGetLastError();
}

cudaError_t launch_fill_chunk_ends(uint64_t *d_ends, size_t n, size_t chunk, unsigned bytes_per_elem,
                                   cudaStream_t stream) {
    if (n == 0 || d_ends == nullptr) return cudaSuccess;
    if (chunk == 0 || chunk > n) chunk = n;
    const size_t n_chunks = (n + chunk - 1) / chunk;
    fill_chunk_ends_kernel<<<dim3(unsigned((n_chunks + 255) / 256)), dim3(256), 0, stream>>>(d_ends, n_chunks, n, chunk,
                                                                                            bytes_per_elem);
    return cudaGetLastError();
}

static const char *kVariantNames[] = {"r4", "r8", "walk", "fused 15x4 d2", "fused 23x2 d2"};
int num_sweep_variants() { return int(sizeof(kVariantNames) / sizeof(kVariantNames[0])); }
const char *sweep_variant_name(int v) { return (v >= 0 && v < num_sweep_variants()) ? kVariantNames[v] : "?"; }

cudaError_t launch_bpe_sweep_pairs(const SweepArgs &a, const uint16_t *d_table, int variant, bool dense_enabled,
                                   cudaStream_t stream, int *host_launches) {
    PairsFE::Params p{d_table};
    // Dense speculation is sound when no even pair can straddle a wall (even chunk size, or one chunk), the
    // output starts on a 16-byte boundary and the output surely fits.
    const size_t chunk = (a.chunk == 0 || a.chunk > a.n) ? a.n : a.chunk;
    const bool single = chunk >= a.n;
    if (dense_enabled && a.n >= 2 && (single || chunk % 2 == 0) && a.out_base_tokens % 8 == 0 &&
        a.out_base_tokens + (a.n + 1) / 2 <= a.out_cap_tokens) {
        int dev = 0;
        cudaError_t err = cudaGetDevice(&dev);
        if (err != cudaSuccess) return err;
        if (dev >= kMaxDevices) return cudaErrorInvalidDevice;
        static std::atomic<bool> configured[kMaxDevices];
        if (!configured[dev].load(std::memory_order_acquire)) {
            err = cudaFuncSetAttribute(dense_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PairsFE::TABLE_BYTES);
            if (err != cudaSuccess) return err;
            configured[dev].store(true, std::memory_order_release);
        }
        // the exact kernels are launched from the device if the speculation fails: they must be configured too
        unsigned exact_grid = 0;
        if (variant == 1) { err = Sweep3Launch<PairsFE, 8>::configure(dev); exact_grid = Sweep3Launch<PairsFE, 8>::grid_for(a.n, dev); }
        else if (variant == 2) { err = Sweep3Launch<PairsFE, 4, true>::configure(dev); exact_grid = Sweep3Launch<PairsFE, 4, true>::grid_for(a.n, dev); }
        else { err = Sweep3Launch<PairsFE, 4>::configure(dev); exact_grid = Sweep3Launch<PairsFE, 4>::grid_for(a.n, dev); }
        if (err != cudaSuccess) return err;
        if (a.scratch.max_tiles < size_t(2 * kMaxRanges)) return cudaErrorInvalidValue;
        // work counter (+256), done counter (+260) and abort word (+384) of the dense pass
        err = cudaMemsetAsync(static_cast<unsigned char *>(a.scratch.ctrl) + 256, 0, 132, stream);
        if (err != cudaSuccess) return err;
        const size_t units = a.n / 16 / kDenseUnitSegs;
        size_t grid = (units + kCtaThreads / 32 - 1) / (kCtaThreads / 32);
        if (grid > size_t(sm_count(dev))) grid = size_t(sm_count(dev));
        if (grid == 0) grid = 1;
        dense_pairs_kernel<<<dim3(unsigned(grid)), dim3(kCtaThreads), PairsFE::TABLE_BYTES, stream>>>(a, d_table, variant,
                                                                                                  exact_grid);
        if (host_launches) *host_launches = kLaunchesDenseAttempt;
        return cudaGetLastError();
    }
    if (variant == 3 && FusedLaunch<15, 4, 2>::applicable(a)) {
        if (host_launches) *host_launches = kLaunchesFused;
        return FusedLaunch<15, 4, 2>::launch(a, d_table, stream);
    }
    if (variant == 4 && FusedLaunch<23, 2, 2>::applicable(a)) {
        if (host_launches) *host_launches = kLaunchesFused;
        return FusedLaunch<23, 2, 2>::launch(a, d_table, stream);
    }
    if (host_launches) *host_launches = kLaunchesExact;
    switch (variant) {
        case 1: return launch_sweep3<PairsFE, 8>(a, p, stream);
        case 2: return launch_sweep3<PairsFE, 4, true>(a, p, stream);
        default: return launch_sweep3<PairsFE, 4>(a, p, stream);
    }
}

cudaError_t launch_bpe_sweep_hash_batch(const SweepArgs *h_args, const SweepArgs *d_args, int nb, const HashTableView &t, bool in_is_u16,
                                        cudaStream_t stream) {
    if (in_is_u16) {
        HashFE<true>::Params p{t};
        return launch_sweep3_batch<HashFE<true>, 8>(h_args, d_args, nb, p, stream);
    }
    ByteMapFE::Params p{t.bytemap};
    return launch_sweep3_batch<ByteMapFE, 4>(h_args, d_args, nb, p, stream);
}

cudaError_t launch_bpe_sweep_hash(const SweepArgs &a, const HashTableView &t, bool in_is_u16, cudaStream_t stream) {
    if (in_is_u16) {
        HashFE<true>::Params p{t};
        return launch_sweep3<HashFE<true>, 8>(a, p, stream);
    }
    ByteMapFE::Params p{t.bytemap};
    return launch_sweep3<ByteMapFE, 4>(a, p, stream);
}

}  // namespace bltk
