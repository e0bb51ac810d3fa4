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
// identity (m all ones) or as a constant (m has a zero).  Carries are resolved with ballots inside
// a warp, a 32-entry table inside a tile and a single-pass decoupled look-back across tiles; the
// same look-back word carries the running output count, so compaction needs no second pass.
#include "kernels.cuh"

#include <atomic>
#include <cstdio>

namespace bltk {
namespace {

constexpr int kCtaThreads = 1024;
constexpr uint32_t FULL = 0xffffffffu;
constexpr int kMaxDevices = 64;
constexpr size_t kCtrlBytes = 64;

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
__device__ __forceinline__ void group_sync(int gid, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
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

__global__ void __launch_bounds__(256) widen_kernel(const uint8_t *__restrict__ in, size_t n,
                                                    uint8_t *__restrict__ out) {
    const size_t nvec = n / 16;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    // two independent 16-byte loads in flight per thread per trip
    for (; i + stride < nvec; i += 2 * stride) {
        const uint4 a = ldg_stream_v4(in + i * 16);
        const uint4 b = ldg_stream_v4(in + (i + stride) * 16);
        stg_v8(out + i * 32, widen8(a.x, a.y), widen8(a.z, a.w));
        stg_v8(out + (i + stride) * 32, widen8(b.x, b.y), widen8(b.z, b.w));
    }
    if (i < nvec) {
        const uint4 a = ldg_stream_v4(in + i * 16);
        stg_v8(out + i * 32, widen8(a.x, a.y), widen8(a.z, a.w));
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

// ================================================================================================
// Look-back descriptors: one 64-bit word per tile, written and read with relaxed 8-byte accesses
// (self-contained, so no fences are needed).
//   bits 63..62  state   0 = not ready, 1 = AGGREGATE, 2 = PREFIX
//   AGGREGATE    bit 61 identity (carry_out = carry_in)   bit 60 constant carry_out (if !identity)
//                bit 59 delta: tokens(carry_in=1) = cnt0 - delta      bits 31..0 cnt0
//   PREFIX       bit 60 carry_out of the tile             bits 55..0 tokens emitted by tiles 0..this
// ================================================================================================
constexpr uint64_t ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62;
constexpr uint64_t F_ID = 1ull << 61, F_CONST = 1ull << 60, F_DELTA = 1ull << 59;
constexpr uint64_t PREFIX_CNT_MASK = (1ull << 56) - 1;

struct Lookback {
    uint32_t carry_in;  // carry entering this tile
    uint64_t base;      // tokens emitted by all earlier tiles of this launch
};

// Executed by one full warp.  tile > 0.
__device__ __forceinline__ Lookback decoupled_lookback(const uint64_t *status, long long tile, int lane) {
    // running = tiles (j+1 .. tile-1) folded into one function of the carry entering tile j+1
    bool run_id = true;
    uint32_t run_const = 0, run_delta = 0;
    uint64_t run_cnt0 = 0;
    long long j = tile - 1;
    for (;;) {
        const long long idx = j - lane;  // lane 0 = nearest predecessor
        uint64_t st;
        uint32_t pmask;
        for (;;) {
            st = (idx >= 0) ? ld_relaxed_u64(status + idx) : ST_PREFIX;  // virtual tile -1: carry 0, count 0
            const uint32_t state = uint32_t(st >> 62);
            pmask = __ballot_sync(FULL, state == 2);
            const uint32_t zmask = __ballot_sync(FULL, state == 0);
            const uint32_t low = pmask & (0u - pmask);                  // nearest PREFIX lane (one-hot)
            const uint32_t need = pmask ? (low | (low - 1)) : FULL;     // lanes 0..p must be ready
            if ((zmask & need) == 0) break;
            __nanosleep(40);
        }
        const int p = pmask ? (__ffs(pmask) - 1) : 32;
        const uint32_t act = (p >= 31) ? FULL : ((2u << p) - 1);        // lanes 0..p
        const bool is_agg = (lane < p);
        const uint32_t idm = __ballot_sync(FULL, is_agg && (st & F_ID));
        const uint32_t nonid = ~idm & act;                              // includes the PREFIX lane
        const uint32_t constm = __ballot_sync(FULL, (st & F_CONST) != 0);
        // carry entering lane's tile = constant of the nearest non-identity lane farther back
        const uint32_t above = nonid & ~((lane == 31) ? FULL : ((2u << lane) - 1));
        const uint32_t cin = above ? ((constm >> (__ffs(above) - 1)) & 1u) : 0u;  // far carry assumed 0
        uint32_t contrib = 0;
        if (is_agg) contrib = uint32_t(st) - ((cin && (st & F_DELTA)) ? 1u : 0u);
        const uint32_t wsum = __reduce_add_sync(FULL, contrib);
        const bool w_id = (nonid == 0);  // only possible without a PREFIX in the window
        const uint32_t w_const = w_id ? 0u : ((constm >> (__ffs(nonid) - 1)) & 1u);
        uint32_t w_delta = 0;
        if (p == 32 && !w_id) {
            const int f = 31 - __clz(nonid);  // farthest non-identity tile sees the unknown far carry
            w_delta = __shfl_sync(FULL, (st & F_DELTA) ? 1u : 0u, f);
        }
        // fold: far = this window, near = running
        const uint32_t c_mid0 = w_id ? 0u : w_const;
        const uint64_t new_cnt0 = uint64_t(wsum) + run_cnt0 - ((c_mid0 && run_delta) ? 1u : 0u);
        const uint32_t new_delta = w_id ? run_delta : w_delta;
        const uint32_t new_const = run_id ? w_const : run_const;
        const bool new_id = w_id && run_id;
        if (p < 32) {
            const uint64_t pcount = __shfl_sync(FULL, st & PREFIX_CNT_MASK, p);
            Lookback r;
            r.carry_in = new_const;  // new_id is false here: the PREFIX lane is a constant
            r.base = pcount + new_cnt0;
            return r;
        }
        run_id = new_id; run_const = new_const; run_delta = new_delta; run_cnt0 = new_cnt0;
        j -= 32;
    }
}

// start bits of one segment: m = pair-membership bits, cin = first element already consumed
__device__ __forceinline__ uint32_t start_bits(uint32_t m, uint32_t cin) {
    const uint32_t mm = m & ~cin;
    const uint32_t s = mm & ~(mm << 1);                      // first bit of every run of ones
    const uint32_t e = mm & ~(mm + (s & 0x55555555u));       // runs that begin at an even position
    return (e & 0x55555555u) | (mm & ~e & 0xAAAAAAAAu);      // same parity as the run's first bit
}

// ================================================================================================
// Front ends: load one 16-byte segment, look every adjacent pair up, return membership bits and
// the big-endian u16 to emit at each position (merged id where the pair is a rule, else the token).
// ================================================================================================

// K2 front end: byte input, direct-indexed byte-pair table in shared memory.
struct PairsFE {
    static constexpr int SEG = 16;       // elements per 16-byte segment
    static constexpr int ELEM = 1;       // bytes per element
    static constexpr int TABLE_BYTES = kPairTableEntries * 2;
    struct Params { const uint16_t *table; };
    const uint16_t *tbl;                 // shared memory

    __device__ __forceinline__ void init(const Params &p, unsigned char *smem) {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.table);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < TABLE_BYTES / 16; i += blockDim.x) dst[i] = src[i];
        tbl = reinterpret_cast<const uint16_t *>(smem);
    }
    __device__ __forceinline__ static uint32_t first_elem(const uint4 &w) { return w.x & 0xffu; }
    __device__ __forceinline__ static uint32_t raw_be(const uint4 &w, int j) {
        const uint32_t word = (j < 4) ? w.x : (j < 8) ? w.y : (j < 12) ? w.z : w.w;
        return ((word >> (8 * (j & 3))) & 0xffu) << 8;  // bswap16(byte)
    }
    __device__ __forceinline__ static uint32_t load_elem(const void *in, size_t pos) {
        return static_cast<const uint8_t *>(in)[pos];
    }
    // vals[8]: 16 big-endian u16, two per register.  Returns membership bits.
    __device__ __forceinline__ uint32_t lookup(const uint4 &w, uint32_t next, uint32_t *vals) const {
        const uint32_t W[5] = {w.x, w.y, w.z, w.w, next};
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int q = j >> 2, k = j & 3;
            uint32_t x;
            if (k < 3) x = __byte_perm(W[q], 0, 0x4400 | ((k + 1) << 4) | k);
            else x = __funnelshift_r(W[q], W[q + 1], 24) & 0xffffu;
            const uint32_t idx = x ^ ((x >> 7) & 0x3Eu);  // == pair_table_index(b0, b1)
            const uint32_t e = tbl[idx];
            m |= ((e & 0xffu) ? 1u : 0u) << j;
            if (j & 1) vals[j >> 1] |= e << 16; else vals[j >> 1] = e;
        }
        return m;
    }
};

// K3 front end: general HashMap<(u16,u16),u16> in global memory (L2-resident), prefiltered by two
// 8 KiB shared-memory bitmaps.  Input is raw bytes (first sweep) or big-endian u16 tokens.
template <bool IN_U16>
struct HashFE {
    static constexpr int SEG = IN_U16 ? 8 : 16;
    static constexpr int ELEM = IN_U16 ? 2 : 1;
    static constexpr int TABLE_BYTES = 2 * 8192;
    struct Params { HashTableView t; };
    const uint32_t *can_left, *can_right;  // shared memory
    const HashSlot *slots;
    uint32_t mask;

    __device__ __forceinline__ void init(const Params &p, unsigned char *smem) {
        uint32_t *s = reinterpret_cast<uint32_t *>(smem);
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
            s[i] = p.t.can_left[i];
            s[2048 + i] = p.t.can_right[i];
        }
        can_left = s;
        can_right = s + 2048;
        slots = p.t.slots;
        mask = p.t.mask;
    }
    // token value (host order) of element j of the segment
    __device__ __forceinline__ static uint32_t elem(const uint4 &w, int j) {
        if (IN_U16) {
            const uint32_t word = (j < 2) ? w.x : (j < 4) ? w.y : (j < 6) ? w.z : w.w;
            const uint32_t be = (word >> (16 * (j & 1))) & 0xffffu;
            return __byte_perm(be, 0, 0x4401);  // bswap16
        } else {
            const uint32_t word = (j < 4) ? w.x : (j < 8) ? w.y : (j < 12) ? w.z : w.w;
            return (word >> (8 * (j & 3))) & 0xffu;
        }
    }
    __device__ __forceinline__ static uint32_t first_elem(const uint4 &w) { return elem(w, 0); }
    __device__ __forceinline__ static uint32_t raw_be(const uint4 &w, int j) {
        return __byte_perm(elem(w, j), 0, 0x4401);
    }
    __device__ __forceinline__ static uint32_t load_elem(const void *in, size_t pos) {
        if (IN_U16) {
            const uint32_t be = static_cast<const uint16_t *>(in)[pos];
            return __byte_perm(be, 0, 0x4401);
        }
        return static_cast<const uint8_t *>(in)[pos];
    }
    __device__ __forceinline__ uint32_t lookup(const uint4 &w, uint32_t next, uint32_t *vals) const {
        uint32_t m = 0;
        uint32_t cur = elem(w, 0);
#pragma unroll
        for (int j = 0; j < SEG; ++j) {
            const uint32_t nxt = (j + 1 < SEG) ? elem(w, (j + 1 < SEG) ? j + 1 : j) : next;
            uint32_t out = cur;
            if (((can_left[cur >> 5] >> (cur & 31)) & (can_right[nxt >> 5] >> (nxt & 31)) & 1u)) {
                const uint32_t key = (cur << 16) | nxt;
                uint32_t h = hash_pair(key) & mask;
                for (;;) {
                    const HashSlot s = slots[h];
                    if (!s.used) break;
                    if (s.key == key) { out = s.value; m |= 1u << j; break; }
                    h = (h + 1) & mask;
                }
            }
            const uint32_t be = __byte_perm(out, 0, 0x4401);
            if (j & 1) vals[j >> 1] |= be << 16; else vals[j >> 1] = be;
            cur = nxt;
        }
        return m;
    }
};

// ================================================================================================
// The sweep kernel.  One persistent CTA per SM (the byte-pair table fills most of shared memory),
// split into GROUPS independent groups of G threads; each group claims tiles of R rounds x G
// segments from an atomic counter (ids are claimed in order by resident groups, so the look-back
// always waits on running work).
// ================================================================================================
struct TileInfo {
    long long tile;            // -1 = no more work
    unsigned long long rem0;   // tile_base % chunk
    unsigned long long ck0;    // tile_base / chunk
};

struct __align__(16) GroupShared {
    uint32_t warp_fn[32];      // per warp-round: bit1 identity, bit0 constant carry_out
    uint32_t warp_cnt[32];     // per warp-round token count (tile carry_in assumed 0)
    TileInfo info[2];
    unsigned long long base;   // look-back result: tokens before this tile
    uint32_t carry_in;         // look-back result
    uint32_t f_idx;            // segment index of the first non-identity segment (or 0xffffffff)
    uint32_t f_delta;          // tokens(carry_in=0) - tokens(carry_in=1) of that segment
    uint32_t total;            // tokens this tile emits (with its real carry_in)
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

template <int G, int R, class FE>
__global__ void __launch_bounds__(kCtaThreads, 1) sweep_kernel(const SweepArgs a, const typename FE::Params fp) {
    constexpr int SEG = FE::SEG;
    constexpr int WARPS = G / 32;
    constexpr int WR = WARPS * R;                 // warp-rounds per tile
    constexpr int ROUND_ELEMS = G * SEG;
    constexpr int TILE_ELEMS = R * ROUND_ELEMS;
    constexpr uint32_t ALL = (SEG == 32) ? FULL : ((1u << SEG) - 1);
    constexpr int STAGE_TOKENS = TILE_ELEMS + 8;
    static_assert(WR <= 32, "a tile holds at most 32 warp-rounds");
    static_assert(SEG % 2 == 0, "identity carry needs an even segment width");

    extern __shared__ __align__(16) unsigned char smem[];
    FE fe;
    fe.init(fp, smem);
    const int gid = threadIdx.x / G;              // group within the CTA
    const int gt = threadIdx.x % G;               // thread within the group
    const int lane = threadIdx.x & 31;
    const int wg = gt >> 5;                       // warp within the group
    unsigned char *gmem = smem + FE::TABLE_BYTES + size_t(gid) * (STAGE_TOKENS * 2 + sizeof(GroupShared));
    uint16_t *stage = reinterpret_cast<uint16_t *>(gmem);
    GroupShared *gs = reinterpret_cast<GroupShared *>(gmem + STAGE_TOKENS * 2);
    const unsigned long long n = a.n;
    const long long n_tiles = (long long)((n + TILE_ELEMS - 1) / TILE_ELEMS);

    auto fetch_tile = [&](TileInfo &ti) {
        const long long t = (long long)atomicAdd(a.scratch.tile_counter, 1u);
        ti.tile = (t < n_tiles) ? t : -1;
        ti.rem0 = 0;
        ti.ck0 = 0;
        if (t < n_tiles && a.chunk != 0) {
            const unsigned long long tb = (unsigned long long)t * TILE_ELEMS;
            ti.ck0 = tb / a.chunk;
            ti.rem0 = tb - ti.ck0 * a.chunk;
        }
    };
    if (gt == 0) fetch_tile(gs->info[0]);
    __syncthreads();  // table + first tile ids visible

    for (int it = 0;; ++it) {
        const TileInfo ti = gs->info[it & 1];
        if (ti.tile < 0) break;
        const unsigned long long tile_base = (unsigned long long)ti.tile * TILE_ELEMS;

        // ---------------- phase A: load, look up, classify ------------------------------------
        uint32_t vals[R][SEG / 2];
        uint32_t mbits[R];      // low SEG bits: m; bit 30: carry_in (tile carry assumed 0); bit 31: depends on tile carry
        uint32_t idb[R], cob[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t off = uint32_t(r * ROUND_ELEMS + gt * SEG);
            const unsigned long long g = tile_base + off;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (g + SEG <= n) {
                w = ldg_stream_v4(static_cast<const unsigned char *>(a.in) + g * FE::ELEM);
            } else if (g < n) {  // ragged last segment: element-wise, never reads past n
                uint32_t tmp[4] = {0, 0, 0, 0};
                for (int j = 0; j < SEG && g + j < n; ++j) {
                    const uint32_t v = FE::load_elem(a.in, g + j);
                    if (FE::ELEM == 1) tmp[j >> 2] |= v << (8 * (j & 3));
                    else tmp[j >> 1] |= __byte_perm(v, 0, 0x4401) << (16 * (j & 1));
                }
                w = make_uint4(tmp[0], tmp[1], tmp[2], tmp[3]);
            }
            uint32_t next = __shfl_down_sync(FULL, FE::first_elem(w), 1);
            if (lane == 31) next = (g + SEG < n) ? FE::load_elem(a.in, g + SEG) : 0u;
            uint32_t m = fe.lookup(w, next, vals[r]);
            const Walls<SEG> wl = seg_walls<SEG, TILE_ELEMS>(a, ti, off, g);
            const uint32_t vm = (g + SEG <= n) ? ALL : (g < n ? ((1u << uint32_t(n - g)) - 1) : 0u);
            m &= vm & ~wl.endm;
            if (wl.endm) {  // a wall suppresses the pair: emit the raw token there, not the merged id
#pragma unroll
                for (int j = 0; j < SEG; ++j) {
                    if ((wl.endm >> j) & 1u) {
                        const uint32_t be = FE::raw_be(w, j);
                        vals[r][j >> 1] = (j & 1) ? ((vals[r][j >> 1] & 0x0000ffffu) | (be << 16))
                                                  : ((vals[r][j >> 1] & 0xffff0000u) | be);
                    }
                }
            }
            const bool id = (m == ALL);
            const uint32_t lead = __clz(~(m << (32 - SEG)));   // ones at the top of the segment
            idb[r] = __ballot_sync(FULL, id);
            cob[r] = __ballot_sync(FULL, (lead & 1u) != 0);
            mbits[r] = m;
            if (lane == 0) {
                const uint32_t nid = ~idb[r];
                const uint32_t wconst = nid ? ((cob[r] >> (31 - __clz(nid))) & 1u) : 0u;
                gs->warp_fn[r * WARPS + wg] = ((nid == 0) ? 2u : 0u) | wconst;
            }
        }
        if (gt == 0) { gs->f_idx = 0xffffffffu; gs->f_delta = 0; }
        group_sync(gid, G);  // #1

        if (gt == 0) fetch_tile(gs->info[(it + 1) & 1]);  // claim the next tile early

        // ---------------- phase B: carries inside the tile, counts -----------------------------
        const uint32_t wfn = (lane < WR) ? gs->warp_fn[lane] : 2u;
        const uint32_t t_idm = __ballot_sync(FULL, (wfn & 2u) != 0);
        const uint32_t t_com = __ballot_sync(FULL, (wfn & 1u) != 0);
        uint32_t excl[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int q = r * WARPS + wg;
            const uint32_t w_nid = ~t_idm & ((1u << q) - 1);
            const bool w_dep = (w_nid == 0);
            const uint32_t w_cin = w_dep ? 0u : ((t_com >> (31 - __clz(w_nid))) & 1u);
            const uint32_t l_nid = ~idb[r] & ((1u << lane) - 1);
            const bool dep = w_dep && (l_nid == 0);
            const uint32_t cin = l_nid ? ((cob[r] >> (31 - __clz(l_nid))) & 1u) : w_cin;
            const uint32_t m = mbits[r];
            const unsigned long long g = tile_base + uint32_t(r * ROUND_ELEMS + gt * SEG);
            const uint32_t vm = (g + SEG <= n) ? ALL : (g < n ? ((1u << uint32_t(n - g)) - 1) : 0u);
            const uint32_t st = start_bits(m, cin);
            const uint32_t cnt = __popc(vm & ~((st << 1) | cin));
            if (dep && m != ALL) {  // the one segment whose count depends on the tile's carry_in
                const uint32_t st1 = start_bits(m, 1u);
                const uint32_t cnt1 = __popc(vm & ~((st1 << 1) | 1u));
                gs->f_idx = uint32_t(r * G + gt);
                gs->f_delta = cnt - cnt1;
            }
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += t;
            }
            excl[r] = incl - cnt;
            if (lane == 31) gs->warp_cnt[q] = incl;
            mbits[r] = m | (cin << 30) | (dep ? (1u << 31) : 0u);
        }
        group_sync(gid, G);  // #2

        // ---------------- phase B2: tile scan, publish, look back -------------------------------
        uint32_t wscan = (lane < WR) ? gs->warp_cnt[lane] : 0u;
        {
            const uint32_t own = wscan;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, wscan, d);
                if (lane >= d) wscan += t;
            }
            const uint32_t total0 = __shfl_sync(FULL, wscan, 31);
            wscan -= own;  // exclusive
            if (wg == 0) {
                const uint32_t t_nid = ~t_idm;
                const bool tile_id = (t_nid == 0);
                const uint32_t tile_const = tile_id ? 0u : ((t_com >> (31 - __clz(t_nid))) & 1u);
                const uint32_t delta = gs->f_delta;
                uint64_t *status = a.scratch.tile_status;
                Lookback lb;
                lb.carry_in = 0;
                lb.base = 0;
                if (ti.tile > 0) {
                    if (lane == 0) {
                        st_relaxed_u64(status + ti.tile, ST_AGG | (tile_id ? F_ID : 0) | (tile_const ? F_CONST : 0) |
                                                             (delta ? F_DELTA : 0) | uint64_t(total0));
                    }
                    lb = decoupled_lookback(status, ti.tile, lane);
                }
                const uint32_t total = total0 - (lb.carry_in ? delta : 0u);
                const uint32_t c_out = tile_id ? lb.carry_in : tile_const;
                if (lane == 0) {
                    st_relaxed_u64(status + ti.tile, ST_PREFIX | (c_out ? F_CONST : 0) | (lb.base + total));
                    gs->base = lb.base;
                    gs->carry_in = lb.carry_in;
                    gs->total = total;
                    if (total < min((unsigned long long)TILE_ELEMS, n - tile_base)) *a.scratch.merged_any = 1u;
                    if (tile_base + TILE_ELEMS >= n) *a.scratch.total_tokens = lb.base + total;
                }
            }
        }
        group_sync(gid, G);  // #3

        // ---------------- phase C: emit into the staging buffer --------------------------------
        const unsigned long long rel_base = gs->base;                   // tokens before this tile (this launch)
        const unsigned long long out_base = rel_base + a.out_base_tokens;  // index into a.out
        const uint32_t tile_cin = gs->carry_in;
        const uint32_t f_idx = gs->f_idx, f_delta = gs->f_delta;
        const uint32_t total = gs->total;
        const uint32_t phase = uint32_t(out_base & 7);                 // keep 16-byte phase of the output
        const bool fits = (out_base + total <= a.out_cap_tokens);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int q = r * WARPS + wg;
            const uint32_t wexcl = __shfl_sync(FULL, wscan, q);
            const uint32_t m = mbits[r] & ALL;
            const bool dep = (mbits[r] >> 31) != 0;
            const uint32_t cin = dep ? tile_cin : ((mbits[r] >> 30) & 1u);
            const uint32_t off = uint32_t(r * ROUND_ELEMS + gt * SEG);
            const unsigned long long g = tile_base + off;
            const uint32_t vm = (g + SEG <= n) ? ALL : (g < n ? ((1u << uint32_t(n - g)) - 1) : 0u);
            const uint32_t st = start_bits(m, cin);
            const uint32_t em = vm & ~((st << 1) | cin);
            const uint32_t seg_idx = uint32_t(r * G + gt);
            uint32_t pos = wexcl + excl[r] - ((tile_cin && seg_idx > f_idx) ? f_delta : 0u);
            if (a.chunk_ends != nullptr) {
                const Walls<SEG> wl = seg_walls<SEG, TILE_ELEMS>(a, ti, off, g);
                uint32_t e = wl.endm;
                unsigned long long ck = wl.ck;
                while (e) {
                    const int d = __ffs(e) - 1;
                    e &= e - 1;
                    const uint32_t upto = __popc(em & ((2u << d) - 1));
                    a.chunk_ends[ck++] = a.chunk_ends_base + 2ull * (rel_base + pos + upto);
                }
            }
            uint32_t sp = phase + pos;
#pragma unroll
            for (int j = 0; j < SEG; ++j) {
                if ((em >> j) & 1u) {
                    const uint32_t v = vals[r][j >> 1];
                    stage[sp++] = uint16_t((j & 1) ? (v >> 16) : v);
                }
            }
        }
        group_sync(gid, G);  // #4

        // ---------------- phase D: staging -> global, 16-byte stores ----------------------------
        if (fits) {
            uint16_t *dst = a.out + (out_base - phase);  // 16-byte aligned
            const uint32_t lo = phase, hi = phase + total;
            for (uint32_t v = gt; v * 8 < hi; v += G) {
                const uint32_t t0 = v * 8;
                if (t0 >= lo && t0 + 8 <= hi) {
                    stg_stream_v4(dst + t0, *reinterpret_cast<const uint4 *>(stage + t0));
                } else {
                    for (uint32_t k = (t0 > lo ? t0 : lo); k < t0 + 8 && k < hi; ++k) dst[k] = stage[k];
                }
            }
        } else if (gt == 0) {
            *a.scratch.overflow = 1u;
        }
    }
}

template <int G, int R, class FE>
constexpr size_t sweep_smem_bytes() {
    return size_t(FE::TABLE_BYTES) + size_t(kCtaThreads / G) * (size_t(R * G * FE::SEG + 8) * 2 + sizeof(GroupShared));
}

template <int G, int R, class FE>
cudaError_t launch_sweep(const SweepArgs &a, const typename FE::Params &fp, cudaStream_t stream) {
    constexpr size_t smem = sweep_smem_bytes<G, R, FE>();
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static_assert(size_t(R) * G * FE::SEG >= kMinTileElems, "status array is sized by kMinTileElems");
    auto kern = sweep_kernel<G, R, FE>;
    static std::atomic<bool> configured[kMaxDevices];  // per instantiation, per device
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev >= kMaxDevices) return cudaErrorInvalidDevice;
    if (!configured[dev].load(std::memory_order_acquire)) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (err != cudaSuccess) return err;
        configured[dev].store(true, std::memory_order_release);
    }
    const int sms = sm_count(dev);
    const size_t tile_elems = size_t(R) * G * FE::SEG;
    const size_t n_tiles = (a.n + tile_elems - 1) / tile_elems;
    if (n_tiles + 1 > a.scratch.max_tiles) return cudaErrorInvalidValue;
    // control block + the descriptors this launch will touch
    err = cudaMemsetAsync(a.scratch.ctrl, 0, kCtrlBytes + n_tiles * 8, stream);
    if (err != cudaSuccess) return err;
    const size_t groups = kCtaThreads / G;
    size_t grid = (n_tiles + groups - 1) / groups;
    if (grid > size_t(sms)) grid = size_t(sms);
    if (grid == 0) grid = 1;
    kern<<<dim3(unsigned(grid)), dim3(kCtaThreads), smem, stream>>>(a, fp);
    return cudaGetLastError();
}

}  // namespace

// ---- scratch -------------------------------------------------------------------------------------
size_t sweep_scratch_bytes(size_t n_elems_max) {
    const size_t tiles = (n_elems_max + kMinTileElems - 1) / kMinTileElems + 1;
    return kCtrlBytes + tiles * 8;
}
SweepScratch sweep_scratch_carve(void *mem, size_t n_elems_max) {
    SweepScratch s;
    const size_t tiles = (n_elems_max + kMinTileElems - 1) / kMinTileElems + 1;
    unsigned char *p = static_cast<unsigned char *>(mem);
    s.ctrl = p;
    s.total_tokens = reinterpret_cast<uint64_t *>(p);
    s.tile_counter = reinterpret_cast<uint32_t *>(p + 8);
    s.merged_any = reinterpret_cast<uint32_t *>(p + 12);
    s.overflow = reinterpret_cast<uint32_t *>(p + 16);
    s.tile_status = reinterpret_cast<uint64_t *>(p + kCtrlBytes);
    s.bytes = kCtrlBytes + tiles * 8;
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
    widen_kernel<<<dim3(unsigned(blocks)), dim3(256), 0, stream>>>(d_in, n, d_out);
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

static const char *kVariantNames[] = {"g256r2", "g512r1", "g512r2", "g1024r1", "g256r1", "g128r2"};
int num_sweep_variants() { return int(sizeof(kVariantNames) / sizeof(kVariantNames[0])); }
const char *sweep_variant_name(int v) { return (v >= 0 && v < num_sweep_variants()) ? kVariantNames[v] : "?"; }

cudaError_t launch_bpe_sweep_pairs(const SweepArgs &a, const uint16_t *d_table, int variant, cudaStream_t stream) {
    PairsFE::Params p{d_table};
    switch (variant) {
        case 1: return launch_sweep<512, 1, PairsFE>(a, p, stream);
        case 2: return launch_sweep<512, 2, PairsFE>(a, p, stream);
        case 3: return launch_sweep<1024, 1, PairsFE>(a, p, stream);
        case 4: return launch_sweep<256, 1, PairsFE>(a, p, stream);
        case 5: return launch_sweep<128, 2, PairsFE>(a, p, stream);
        default: return launch_sweep<256, 2, PairsFE>(a, p, stream);
    }
}

cudaError_t launch_bpe_sweep_hash(const SweepArgs &a, const HashTableView &t, bool in_is_u16, cudaStream_t stream) {
    if (in_is_u16) {
        HashFE<true>::Params p{t};
        return launch_sweep<256, 2, HashFE<true>>(a, p, stream);
    }
    HashFE<false>::Params p{t};
    return launch_sweep<256, 2, HashFE<false>>(a, p, stream);
}

}  // namespace bltk
