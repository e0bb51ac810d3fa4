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

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace bltk {
namespace {

constexpr int kCtaThreads = 1024;
constexpr uint32_t FULL = 0xffffffffu;
constexpr int kMaxDevices = 64;
constexpr size_t kCtrlBytes = 512;  // four 128-byte lines: tile counter | phase hint | results | spare

int sm_count(int dev) {
    static std::atomic<int> cached[kMaxDevices];
    int v = cached[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cached[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// ---- optional in-kernel trace (debug builds only: -DBLT_TRACE) --------------------------------------
#ifdef BLT_TRACE
__device__ unsigned long long *g_trace = nullptr;   // [cta][iter][8] globaltimer stamps of group 0, lane 0 of warp 0
__device__ unsigned int g_trace_iters = 0;
__device__ unsigned long long g_dbg_word = 0;  // written by the look-back of group 0 / CTA 0 only (racy, debug)
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TRACE(ev)                                                                                   \
    do {                                                                                            \
        if (g_trace && threadIdx.x == 0 && trace_it < g_trace_iters)                                \
            g_trace[(size_t(blockIdx.x) * g_trace_iters + trace_it) * 8 + (ev)] = gtime();          \
    } while (0)
#else
#define TRACE(ev) do { } while (0)
#endif

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
// (self-contained, so no fences are needed).  A tile's effect on the scan is a function of the carry
// entering it: carry_in in {0,1} -> (carry_out, tokens emitted).  A tile may know only one branch of
// that function (see "speculated phase" below), so every branch has its own valid bit.
//   bits 63..62  state   0 = not ready, 1 = AGGREGATE, 2 = PREFIX
//   AGGREGATE    bit 61 v0, bit 60 v1 (branch known)   bit 59 c0, bit 58 c1 (carry_out per branch)
//                bits 23..0 tokens if carry_in = 0      bits 47..24 tokens if carry_in = 1
//   PREFIX       bit 59 carry_out of the tile           bits 55..0 tokens emitted by tiles 0..this
// ================================================================================================
constexpr uint64_t ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62;
constexpr uint64_t PREFIX_CNT_MASK = (1ull << 56) - 1;

// f: bit0 v0, bit1 v1, bit2 c0, bit3 c1
struct Fn {
    uint32_t f, cnt0, cnt1;
};
constexpr uint32_t FN_IDENTITY = 0xBu;  // both branches known, carry passes through, no tokens

__device__ __forceinline__ Fn fn_compose(const Fn &far, const Fn &near) {  // carry flows far -> near
    const uint32_t m0 = (far.f >> 2) & 1u, m1 = (far.f >> 3) & 1u;
    Fn r;
    r.f = (far.f & (near.f >> m0) & 1u) | (((far.f >> 1) & (near.f >> m1) & 1u) << 1) |
          (((near.f >> (2 + m0)) & 1u) << 2) | (((near.f >> (2 + m1)) & 1u) << 3);
    r.cnt0 = far.cnt0 + (m0 ? near.cnt1 : near.cnt0);
    r.cnt1 = far.cnt1 + (m1 ? near.cnt1 : near.cnt0);
    return r;
}
__device__ __forceinline__ Fn fn_decode(uint64_t st) {
    Fn r;
    if ((st >> 62) == 2) {  // PREFIX acts as a constant; its absolute count is handled separately
        const uint32_t c = uint32_t(st >> 59) & 1u;
        r.f = 3u | (c << 2) | (c << 3);
        r.cnt0 = r.cnt1 = 0;
    } else {
        r.f = (uint32_t(st >> 61) & 1u) | ((uint32_t(st >> 60) & 1u) << 1) | ((uint32_t(st >> 59) & 1u) << 2) |
              ((uint32_t(st >> 58) & 1u) << 3);
        r.cnt0 = uint32_t(st) & 0xFFFFFFu;
        r.cnt1 = uint32_t(st >> 24) & 0xFFFFFFu;
    }
    return r;
}
__device__ __forceinline__ uint64_t agg_encode(uint32_t f, uint32_t cnt0, uint32_t cnt1) {
    return ST_AGG | (uint64_t(f & 1u) << 61) | (uint64_t((f >> 1) & 1u) << 60) | (uint64_t((f >> 2) & 1u) << 59) |
           (uint64_t((f >> 3) & 1u) << 58) | (uint64_t(cnt1 & 0xFFFFFFu) << 24) | uint64_t(cnt0 & 0xFFFFFFu);
}

// Executed by one full warp for tile > 0.  Every lane polls LW consecutive predecessors, so one
// round trip to L2 covers 32*LW tiles.  Returns false if the branch it needs of some predecessor is
// not known yet (that tile is upgrading).
//
// Fast fold: under the hypothesis "the carry is `hyp` all the way" every entry between the nearest
// PREFIX and this tile only has to confirm that its branch `hyp` is known and hands `hyp` on; the
// token counts of that branch are then simply summed.  That is the steady state of merge-dense input
// (every tile PART(hyp)) and costs a handful of instructions per entry.  Anything else falls back to
// the general fold, which composes the partial functions entry by entry.
template <int LW>
__device__ __noinline__ bool decoupled_lookback(const uint64_t *status, long long tile, int lane, uint32_t hyp,
                                                uint32_t *carry_in, uint64_t *base) {
#ifdef BLT_TRACE
    unsigned int dbg_polls = 0, dbg_retries = 0, dbg_general = 0;
#define DBG_DONE() do { if (lane == 0) g_dbg_word = (unsigned long long)dbg_polls | ((unsigned long long)dbg_retries << 16) | ((unsigned long long)dbg_general << 32); } while (0)
#else
#define DBG_DONE() do { } while (0)
#endif
    // ---------------- fast fold ----------------
    // Row-major polling: row i, lane l reads tile j - (32*i + l), so every row is one coalesced 256-byte
    // request (entry 0 = nearest predecessor).
    {
        uint64_t sum = 0;
        long long j = tile - 1;
        for (;;) {
            uint64_t st[LW];
            uint32_t pmask[LW], zmask[LW];
#pragma unroll
            for (int i = 0; i < LW; ++i) {
                const long long idx = j - (long long)(i * 32 + lane);
                st[i] = (idx >= 0) ? ld_relaxed_u64(status + idx) : ST_PREFIX;  // virtual tile -1
            }
            int ip = LW;          // row of the nearest PREFIX
            uint32_t below = FULL;  // lanes of that row that are nearer than the PREFIX
            bool ready = true;
#pragma unroll
            for (int i = 0; i < LW; ++i) {
                const uint32_t state = uint32_t(st[i] >> 62);
                pmask[i] = __ballot_sync(FULL, state == 2);
                zmask[i] = __ballot_sync(FULL, state == 0);
                if (ip == LW) {
                    if (pmask[i]) {
                        ip = i;
                        below = (pmask[i] & (0u - pmask[i])) - 1u;
                        ready = ready && ((zmask[i] & below) == 0);
                    } else {
                        ready = ready && (zmask[i] == 0);
                    }
                }
            }
#ifdef BLT_TRACE
            ++dbg_polls;
#endif
            if (!ready) {  // somebody nearer than the nearest PREFIX has not published yet: poll again
#ifdef BLT_TRACE
                ++dbg_retries;
#endif
                continue;
            }
            uint32_t part_sum = 0;
            bool good = true;
            uint64_t pword = 0;
#pragma unroll
            for (int i = 0; i < LW; ++i) {
                const bool nearer = (i < ip) || (i == ip && ((below >> lane) & 1u));
                if (nearer) {
                    const uint32_t hi = uint32_t(st[i] >> 32);
                    // branch `hyp` known (bit 61 - hyp) and its carry_out (bit 59 - hyp) equals hyp
                    const uint32_t known = (hi >> (29 - hyp)) & 1u;
                    const uint32_t cout = (hi >> (27 - hyp)) & 1u;
                    good = good && known && (cout == hyp);
                    part_sum += uint32_t(st[i] >> (24 * hyp)) & 0xFFFFFFu;
                }
                if (i == ip && ((below + 1u) >> lane) == 1u) pword = st[i];  // the PREFIX lane itself
            }
            if (!__all_sync(FULL, good)) break;  // not a pure `hyp` chain: general fold
            sum += __reduce_add_sync(FULL, part_sum);
            if (ip < LW) {
                const int plane = __ffs(below + 1u) - 1;
                pword = __shfl_sync(FULL, pword, plane);
                if ((uint32_t(pword >> 59) & 1u) != hyp) break;  // chain is fine but starts from the other carry
                *carry_in = hyp;
                *base = (pword & PREFIX_CNT_MASK) + sum;
                DBG_DONE();
                return true;
            }
            j -= 32 * LW;
        }
    }
    // ---------------- general fold ----------------
#ifdef BLT_TRACE
    dbg_general = 1;
    DBG_DONE();
#endif
    uint32_t run_f = FN_IDENTITY;  // tiles (j+1 .. tile-1) as a function of the carry entering tile j+1
    uint64_t run_c0 = 0, run_c1 = 0;
    long long j = tile - 1;
    for (;;) {
        uint64_t st[LW];
        int p_lane, p_sub;
        for (;;) {
            int lp = LW, lz = LW;
#pragma unroll
            for (int i = LW - 1; i >= 0; --i) {
                const long long idx = j - (long long)(lane * LW + i);
                st[i] = (idx >= 0) ? ld_relaxed_u64(status + idx) : ST_PREFIX;
                const uint32_t state = uint32_t(st[i] >> 62);
                if (state == 2) lp = i;
                if (state == 0) lz = i;
            }
            const uint32_t pm = __ballot_sync(FULL, lp < LW);
            p_lane = pm ? (__ffs(pm) - 1) : 32;
            p_sub = pm ? __shfl_sync(FULL, lp, p_lane & 31) : LW;
            const bool bad = (lane < p_lane && lz < LW) || (lane == p_lane && lz < p_sub);
            if (!__any_sync(FULL, bad)) break;
            __nanosleep(20);
        }
        Fn acc;
        acc.f = FN_IDENTITY;
        acc.cnt0 = acc.cnt1 = 0;
        uint64_t pcount = 0;
#pragma unroll
        for (int i = LW - 1; i >= 0; --i) {  // far -> near inside the lane
            const bool beyond = (lane > p_lane) || (lane == p_lane && i > p_sub);
            if (!beyond) acc = fn_compose(acc, fn_decode(st[i]));
            if (lane == p_lane && i == p_sub) pcount = st[i] & PREFIX_CNT_MASK;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {  // lane l+d is farther back than lane l
            Fn o;
            o.f = __shfl_down_sync(FULL, acc.f, d);
            o.cnt0 = __shfl_down_sync(FULL, acc.cnt0, d);
            o.cnt1 = __shfl_down_sync(FULL, acc.cnt1, d);
            if (lane + d < 32) acc = fn_compose(o, acc);
        }
        const uint32_t wf = __shfl_sync(FULL, acc.f, 0);
        const uint32_t wc0 = __shfl_sync(FULL, acc.cnt0, 0), wc1 = __shfl_sync(FULL, acc.cnt1, 0);
        // total = compose(far = this window, near = running), 64-bit counts
        const uint32_t m0 = (wf >> 2) & 1u, m1 = (wf >> 3) & 1u;
        const uint32_t nf = (wf & (run_f >> m0) & 1u) | (((wf >> 1) & (run_f >> m1) & 1u) << 1) |
                            (((run_f >> (2 + m0)) & 1u) << 2) | (((run_f >> (2 + m1)) & 1u) << 3);
        const uint64_t n0 = uint64_t(wc0) + (m0 ? run_c1 : run_c0);
        const uint64_t n1 = uint64_t(wc1) + (m1 ? run_c1 : run_c0);
        run_f = nf; run_c0 = n0; run_c1 = n1;
        if (p_lane < 32) {
            pcount = __shfl_sync(FULL, pcount, p_lane);
            if (!(run_f & 1u)) return false;  // a PREFIX is constant: both branches agree, test branch 0
            *carry_in = (run_f >> 2) & 1u;
            *base = pcount + run_c0;
            return true;
        }
        j -= 32 * LW;
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
    __device__ __forceinline__ static uint32_t present_mask(const uint32_t *vals) {
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) m |= (((vals[i >> 1] >> (16 * (i & 1))) & 0xffu) ? 1u : 0u) << i;
        return m;
    }
    __device__ __forceinline__ uint32_t lookup_half(const uint4 &w, uint32_t next, uint32_t par, uint32_t *vals) const {
        lookup_vals(w, next, par, vals);
        return present_mask(vals);
    }
};

// K3 front end: general HashMap<(u16,u16),u16> in global memory (L2-resident), prefiltered by two
// 8 KiB shared-memory bitmaps.  Input is raw bytes (first sweep) or big-endian u16 tokens.
template <bool IN_U16>
struct HashFE {
    static constexpr int SEG = IN_U16 ? 8 : 16;
    static constexpr int ELEM = IN_U16 ? 2 : 1;
    static constexpr int TABLE_BYTES = 2 * 8192;
    static constexpr bool kMembershipInValue = false;  // ids may be < 256 and may equal the element
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
        if (((can_left[cur >> 5] >> (cur & 31)) & (can_right[nxt >> 5] >> (nxt & 31)) & 1u)) {
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

// ================================================================================================
// K2-dense: the speculative streaming form of the sweep for merge-dense input.
//
// If every pair that starts at an EVEN offset of its chunk is a rule, the reference's scan merges
// exactly those pairs (start[0] = 1, start[1] = 0, start[2] = 1, ...), whatever the odd pairs are: the
// output is the looked-up id of every even pair, token k of the launch sits at out[k], and no carry
// or count has to cross a tile.  With an even chunk size no such pair straddles a wall.  The kernel
// streams under that hypothesis - 16 bytes in, 8 lookups, 16 bytes out per lane, no barrier, no
// look-back - and raises `abort` at the first pair that is not a rule.  The exact sweep kernel is
// always enqueued behind it and returns at once when the hypothesis held.
// ================================================================================================
__global__ void __launch_bounds__(kCtaThreads, 1)
dense_pairs_kernel(const unsigned char *__restrict__ in, unsigned long long n, uint16_t *__restrict__ out,
                   const uint16_t *__restrict__ table, uint32_t *abort_flag) {
    extern __shared__ __align__(16) unsigned char smem[];
    PairsFE fe;
    PairsFE::Params p{table};
    fe.init(p, smem);
    __syncthreads();
    const unsigned long long n_segs = n / 16;
    const unsigned long long stride = (unsigned long long)gridDim.x * kCtaThreads;
    unsigned long long seg = (unsigned long long)blockIdx.x * kCtaThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool bad = false;
    uint32_t ab = 0;  // the abort word as it was one trip ago: never waited for inside a trip
    constexpr int U = 4;  // independent 16-byte loads in flight per thread
    for (; seg + (U - 1) * stride < n_segs; seg += U * stride) {
        uint4 w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = ldg_stream_v4(in + (seg + u * stride) * 16);
        const uint32_t ab_now = ab;
        if (lane == 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(ab) : "l"(abort_flag) : "memory");
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t v[4];
            fe.lookup_vals(w[u], 0u, 0u, v);
            bad = bad || !PairsFE::all_present(v);
            stg_stream_v4(out + (seg + u * stride) * 8, make_uint4(v[0], v[1], v[2], v[3]));
        }
        if (__any_sync(FULL, bad || ab_now != 0)) break;  // this warp or somebody else gave up: stop streaming
    }
    if (!__any_sync(FULL, bad || ab != 0)) {
        for (; seg < n_segs; seg += stride) {
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
    if (bad) *reinterpret_cast<volatile uint32_t *>(abort_flag) = 1u;
}

// ================================================================================================
// The sweep kernel.  One persistent CTA per SM (the byte-pair table fills most of shared memory),
// split into GROUPS independent groups of G threads.  Tiles (R rounds x G segments) are assigned
// statically, tile = iteration * total_groups + group; the kernel is launched cooperatively so that
// every group is resident, which is what lets a group spin on its predecessors' descriptors.
//
// Speculated phase.  In merge-dense input nearly every adjacent pair is a rule, so the starts are
// simply the positions of one parity p (the carry entering the tile) and the pairs of the other
// parity never matter.  A tile therefore first looks up only the pairs of the parity it PREDICTS
// (0 at a chunk start, else the carry_out of the tile this group finished last).  If all of them are
// rules the tile is "PART": it knows branch p of its function (carry_out = p, exactly half as many
// tokens as elements), publishes that, and if the look-back confirms carry_in = p it stores the
// looked-up ids straight from registers with 16-byte stores.  Otherwise (a miss, a wall in the way, or
// a wrong prediction) it looks up the other parity as well and takes the general path: full
// membership bits, run parity, scan, staged compaction.
// ================================================================================================
struct TileInfo {
    unsigned long long rem0;   // tile_base % chunk (tile_base itself when there are no walls)
    unsigned long long ck0;    // tile_base / chunk
};

enum : uint32_t { ACT_GO = 0, ACT_FULL = 1 };

struct __align__(16) GroupShared {
    uint32_t warp_ok[32];      // per warp (A1): every lane found all pairs of the predicted parity
    uint32_t warp_fn[64];      // per warp-round (A2): bit1 identity, bit0 constant carry_out
    uint32_t warp_cnt[64];     // per warp-round token count (tile carry_in assumed 0)
    unsigned long long base;   // look-back result: tokens before this tile
    uint32_t carry_in;         // look-back result
    uint32_t carry_out;        // of this tile, once known
    uint32_t f_idx;            // segment index of the first non-identity segment (or 0xffffffff)
    uint32_t f_delta;          // tokens(carry_in=0) - tokens(carry_in=1) of that segment
    uint32_t total;            // tokens this tile emits (with its real carry_in)
    uint32_t action;           // ACT_*
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

__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int G, int R, class FE>
__global__ void __launch_bounds__(kCtaThreads, 1) sweep_kernel(const SweepArgs a, const typename FE::Params fp) {
    constexpr int SEG = FE::SEG;
    constexpr int HV = SEG / 4;                   // registers holding one parity's SEG/2 tokens
    constexpr int GROUPS = kCtaThreads / G;
    constexpr int WARPS = G / 32;
    constexpr int WR = WARPS * R;                 // warp-rounds per tile
    constexpr int ROUND_ELEMS = G * SEG;
    constexpr int TILE_ELEMS = R * ROUND_ELEMS;
    constexpr uint32_t ALL = (1u << SEG) - 1;
    constexpr uint32_t EVEN = 0x55555555u & ALL;
    constexpr uint32_t HALF_ALL = (1u << (SEG / 2)) - 1;
    constexpr int STAGE_TOKENS = TILE_ELEMS + 8;
    constexpr int LW = 8;                         // look-back: 256 predecessors per poll
    static_assert(WR <= 64, "a tile holds at most 64 warp-rounds");
    static_assert(SEG == 16 || SEG == 8, "segment is one 16-byte vector");
    static_assert(2 * GROUPS + 1 <= 16, "two named barriers per group");

    if (a.dense_flag != nullptr && *reinterpret_cast<const volatile uint32_t *>(a.dense_flag) == 0u) {
        // The dense pass in front of this launch already produced the whole output: publish its totals.
        if (blockIdx.x == 0) {
            const unsigned long long tokens = (a.n + 1) / 2;
            if (threadIdx.x == 0) {
                *a.scratch.total_tokens = tokens;
                *a.scratch.merged_any = (a.n >= 2) ? 1u : 0u;
            }
            if (a.chunk_ends != nullptr) {
                const unsigned long long c = (a.chunk == 0 || a.chunk > a.n) ? a.n : a.chunk;
                const unsigned long long n_chunks = (a.n + c - 1) / c;
                for (unsigned long long k = threadIdx.x; k < n_chunks; k += blockDim.x)
                    a.chunk_ends[k] = a.chunk_ends_base + ((k + 1 == n_chunks) ? 2 * tokens : (k + 1) * c);
            }
        }
        return;
    }
    extern __shared__ __align__(16) unsigned char smem[];
    FE fe;
    fe.init(fp, smem);
    const int gid = threadIdx.x / G;              // group within the CTA
    const int gt = threadIdx.x % G;               // thread within the group
    const int lane = threadIdx.x & 31;
    const int wg = gt >> 5;                       // warp within the group
    const int bar_a = 1 + 2 * gid, bar_b = 2 + 2 * gid;
    unsigned char *gmem = smem + FE::TABLE_BYTES + size_t(gid) * (STAGE_TOKENS * 2 + sizeof(GroupShared));
    uint16_t *stage = reinterpret_cast<uint16_t *>(gmem);
    GroupShared *gs = reinterpret_cast<GroupShared *>(gmem + STAGE_TOKENS * 2);
    const unsigned long long n = a.n;
    const long long n_tiles = (long long)((n + TILE_ELEMS - 1) / TILE_ELEMS);
    uint64_t *const status = a.scratch.tile_status;
    const long long total_groups = (long long)gridDim.x * GROUPS;
    const unsigned long long stride_elems = (unsigned long long)total_groups * TILE_ELEMS;

    // loads the R segments of a tile (full vectors only; a ragged last segment is fetched element-wise
    // later) and, in lane 31, the look-ahead element of every segment
    auto load_tile = [&](long long tile, uint4 *w, uint32_t *nx) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned long long g = (unsigned long long)tile * TILE_ELEMS + uint32_t(r * ROUND_ELEMS + gt * SEG);
            w[r] = make_uint4(0, 0, 0, 0);
            nx[r] = 0;
            if (tile < n_tiles && g + SEG <= n) w[r] = ldg_stream_v4(static_cast<const unsigned char *>(a.in) + g * FE::ELEM);
            if (lane == 31 && tile < n_tiles && g + SEG < n) nx[r] = FE::load_elem(a.in, g + SEG);
        }
    };

    long long tile = (long long)blockIdx.x * GROUPS + gid;
    TileInfo ti;
    ti.rem0 = (unsigned long long)tile * TILE_ELEMS;
    ti.ck0 = 0;
    if (a.chunk != 0) { ti.ck0 = ti.rem0 / a.chunk; ti.rem0 -= ti.ck0 * a.chunk; }
    uint32_t phat_next = 0;   // carry_out of the tile this group finished last
    bool merged = false;
    uint4 w[R];
    uint32_t nx31[R];
    load_tile(tile, w, nx31);
    __syncthreads();  // table visible

    unsigned int trace_it = 0;
    (void)trace_it;
    for (; tile < n_tiles; tile += total_groups, ++trace_it) {
        TRACE(0);
        const unsigned long long tile_base = (unsigned long long)tile * TILE_ELEMS;
        // a chunk start has carry 0 by definition; elsewhere dense runs keep their phase
        const uint32_t phat = (ti.rem0 == 0) ? 0u : phat_next;
        const uint32_t par_mask = (EVEN << phat) & ALL;   // positions whose pairs phase `phat` merges
        // tile-uniform classification: nothing special inside the tile (no end of input, no wall except
        // possibly at the very last position, which an even phase never looks at)
        const bool end_wall = (a.chunk != 0) && (ti.rem0 + TILE_ELEMS == a.chunk);
        const bool lean = FE::kMembershipInValue && (tile_base + TILE_ELEMS < n) &&
                          (a.chunk == 0 || ti.rem0 + TILE_ELEMS <= a.chunk) && !(end_wall && phat);

        // ---------------- phase A1: look up the pairs of the predicted parity ---------------------
        uint32_t hv[R][HV];     // tokens of parity phat (after a swap in A2: of the even positions)
        uint32_t ov[R][HV];     // tokens of the other parity (A2 only)
        uint32_t pres[R];       // membership bits of parity phat (front ends without kMembershipInValue)
        bool lane_ok = true;
        if (lean) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                uint32_t next = __shfl_down_sync(FULL, FE::first_elem(w[r]), 1);
                if (lane == 31) next = nx31[r];
                nx31[r] = next;  // from here on: this lane's look-ahead element
                fe.lookup_vals(w[r], next, phat, hv[r]);
                lane_ok = lane_ok && FE::all_present(hv[r]);
                pres[r] = 0;
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t off = uint32_t(r * ROUND_ELEMS + gt * SEG);
                const unsigned long long g = tile_base + off;
                if (g < n && g + SEG > n) {  // ragged last segment: element-wise, never reads past n
                    uint32_t tmp[4] = {0, 0, 0, 0};
                    for (int j = 0; j < SEG && g + j < n; ++j) {
                        const uint32_t v = FE::load_elem(a.in, g + j);
                        if (FE::ELEM == 1) tmp[j >> 2] |= v << (8 * (j & 3));
                        else tmp[j >> 1] |= __byte_perm(v, 0, 0x4401) << (16 * (j & 1));
                    }
                    w[r] = make_uint4(tmp[0], tmp[1], tmp[2], tmp[3]);
                }
                uint32_t next = __shfl_down_sync(FULL, FE::first_elem(w[r]), 1);
                if (lane == 31) next = nx31[r];
                nx31[r] = next;
                const Walls<SEG> wl = seg_walls<SEG, TILE_ELEMS>(a, ti, off, g);
                const uint32_t vm = (g + SEG <= n) ? ALL : (g < n ? ((1u << uint32_t(n - g)) - 1) : 0u);
                const uint32_t allow = vm & ~wl.endm;
                pres[r] = fe.lookup_half(w[r], next, phat, hv[r]);
                lane_ok = lane_ok && (pres[r] == HALF_ALL) && ((allow & par_mask) == par_mask);
            }
        }
        {
            const uint32_t okb = __ballot_sync(FULL, lane_ok);
            if (lane == 0) gs->warp_ok[wg] = (okb == FULL) ? 1u : 0u;
        }
        TRACE(1);
        // barrier A: warp 0 waits for every warp's flag, the other warps only signal and go on to wait
        // for warp 0's verdict at barrier B
        if (wg == 0) {
            bar_sync(bar_a, G);
            TRACE(2);
            const bool part = __all_sync(FULL, (lane < WARPS) ? (gs->warp_ok[lane] != 0) : true);
            uint32_t cin = 0;
            uint64_t base = 0;
            bool ok = part;
            if (part && tile > 0) {
                if (lane == 0) {
                    const uint32_t f = (1u << phat) | (phat << (2 + phat));
                    st_relaxed_u64(status + tile, agg_encode(f, phat ? 0u : TILE_ELEMS / 2, phat ? TILE_ELEMS / 2 : 0u));
                }
                TRACE(3);
                ok = decoupled_lookback<LW>(status, tile, lane, phat, &cin, &base);
                TRACE(4);
#ifdef BLT_TRACE
                if (g_trace && threadIdx.x == 0 && trace_it < g_trace_iters)
                    g_trace[(size_t(blockIdx.x) * g_trace_iters + trace_it) * 8 + 7] = g_dbg_word;
#endif
            }
            if (ok && cin == phat) {
                if (lane == 0) {
                    st_relaxed_u64(status + tile, ST_PREFIX | (uint64_t(phat) << 59) | (base + TILE_ELEMS / 2));
                    gs->base = base;
                    gs->carry_in = cin;
                    gs->carry_out = phat;
                    gs->total = TILE_ELEMS / 2;
                    gs->action = ACT_GO;
                    if (tile_base + TILE_ELEMS >= n) *a.scratch.total_tokens = base + TILE_ELEMS / 2;
                }
            } else if (lane == 0) {
                gs->action = ACT_FULL;  // a miss, a wrong phase, or somebody before us has to upgrade first
            }
        } else {
            bar_arrive(bar_a, G);
        }
        bar_sync(bar_b, G);  // #B
        TRACE(5);
        const bool part = (gs->action == ACT_GO);
        uint32_t mbits[R], idb[R], cob[R], excl[R], allow[R];
        uint32_t wscan0 = 0, wscan1 = 0;  // exclusive scan of warp-round counts (entries lane, lane+32)

        if (!part) {
            // ------------ phase A2: the other parity, full membership, segment functions ----------
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t off = uint32_t(r * ROUND_ELEMS + gt * SEG);
                const unsigned long long g = tile_base + off;
                const Walls<SEG> wl = seg_walls<SEG, TILE_ELEMS>(a, ti, off, g);
                const uint32_t vm = (g + SEG <= n) ? ALL : (g < n ? ((1u << uint32_t(n - g)) - 1) : 0u);
                allow[r] = vm & ~wl.endm;
                uint32_t hp = FE::kMembershipInValue ? FE::present_mask(hv[r]) : pres[r];
                uint32_t op = fe.lookup_half(w[r], nx31[r], phat ^ 1u, ov[r]);
                if (phat) {  // name the arrays by position parity: hv = even positions, ov = odd
#pragma unroll
                    for (int k = 0; k < HV; ++k) { const uint32_t t = hv[r][k]; hv[r][k] = ov[r][k]; ov[r][k] = t; }
                    const uint32_t t = hp; hp = op; op = t;
                }
                const uint32_t m = (spread_even(hp) | (spread_even(op) << 1)) & allow[r] & ALL;
                if (wl.endm) {  // a wall suppresses the pair: emit the raw token there, not the merged id
#pragma unroll
                    for (int j = 0; j < SEG; ++j) {
                        if ((wl.endm >> j) & 1u) {
                            const uint32_t be = FE::raw_be(w[r], j);
                            uint32_t &dst = (j & 1) ? ov[r][j >> 2] : hv[r][j >> 2];
                            dst = ((j >> 1) & 1) ? ((dst & 0x0000ffffu) | (be << 16)) : ((dst & 0xffff0000u) | be);
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
            bar_sync(bar_a, G);  // #1b

            // ------------ phase B: carries inside the tile (tile carry_in assumed 0), counts --------
            const uint32_t f0 = (lane < WR) ? gs->warp_fn[lane] : 2u;
            const uint32_t f1 = (lane + 32 < WR) ? gs->warp_fn[lane + 32] : 2u;
            const unsigned long long t_idm = (unsigned long long)__ballot_sync(FULL, (f0 & 2u) != 0) |
                                             ((unsigned long long)__ballot_sync(FULL, (f1 & 2u) != 0) << 32);
            const unsigned long long t_com = (unsigned long long)__ballot_sync(FULL, (f0 & 1u) != 0) |
                                             ((unsigned long long)__ballot_sync(FULL, (f1 & 1u) != 0) << 32);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int q = r * WARPS + wg;
                const unsigned long long w_nid = ~t_idm & ((1ull << q) - 1);
                const bool w_dep = (w_nid == 0);
                const uint32_t w_cin = w_dep ? 0u : uint32_t((t_com >> (63 - __clzll((long long)w_nid))) & 1ull);
                const uint32_t l_nid = ~idb[r] & ((1u << lane) - 1);
                const bool dep = w_dep && (l_nid == 0);
                const uint32_t cin = l_nid ? ((cob[r] >> (31 - __clz(l_nid))) & 1u) : w_cin;
                const uint32_t m = mbits[r];
                const unsigned long long g = tile_base + uint32_t(r * ROUND_ELEMS + gt * SEG);
                const uint32_t valid = (g + SEG <= n) ? ALL : (g < n ? ((1u << uint32_t(n - g)) - 1) : 0u);
                const uint32_t st = start_bits(m, cin);
                const uint32_t cnt = __popc(valid & ~((st << 1) | cin));
                if (dep && m != ALL) {  // the one segment whose count depends on the tile's carry_in
                    const uint32_t st1 = start_bits(m, 1u);
                    const uint32_t cnt1 = __popc(valid & ~((st1 << 1) | 1u));
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
            bar_sync(bar_b, G);  // #2
            // tile scan over up to 64 warp-round counts (every warp redundantly)
            {
                const uint32_t c0 = (lane < WR) ? gs->warp_cnt[lane] : 0u;
                const uint32_t c1 = (lane + 32 < WR) ? gs->warp_cnt[lane + 32] : 0u;
                uint32_t i0 = c0, i1 = c1;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t0 = __shfl_up_sync(FULL, i0, d);
                    const uint32_t t1 = __shfl_up_sync(FULL, i1, d);
                    if (lane >= d) { i0 += t0; i1 += t1; }
                }
                const uint32_t sum0 = __shfl_sync(FULL, i0, 31);
                wscan0 = i0 - c0;
                wscan1 = sum0 + i1 - c1;
                if (wg == 0) {
                    const uint32_t total0 = sum0 + __shfl_sync(FULL, i1, 31);
                    const unsigned long long t_nid = ~t_idm;
                    const bool tile_id = (t_nid == 0);
                    const uint32_t tile_const = tile_id ? 0u : uint32_t((t_com >> (63 - __clzll((long long)t_nid))) & 1ull);
                    const uint32_t delta = gs->f_delta;
                    uint32_t cin = 0;
                    uint64_t base = 0;
                    if (tile > 0) {
                        const uint32_t f = 3u | ((tile_id ? 0u : tile_const) << 2) | ((tile_id ? 1u : tile_const) << 3);
                        if (lane == 0) st_relaxed_u64(status + tile, agg_encode(f, total0, total0 - delta));
                        while (!decoupled_lookback<LW>(status, tile, lane, phat, &cin, &base)) __nanosleep(100);
                    }
                    const uint32_t total = total0 - (cin ? delta : 0u);
                    const uint32_t c_out = tile_id ? cin : tile_const;
                    if (lane == 0) {
                        st_relaxed_u64(status + tile, ST_PREFIX | (uint64_t(c_out) << 59) | (base + total));
                        gs->base = base;
                        gs->carry_in = cin;
                        gs->carry_out = c_out;
                        gs->total = total;
                        if (total < min((unsigned long long)TILE_ELEMS, n - tile_base)) merged = true;
                        if (tile_base + TILE_ELEMS >= n) *a.scratch.total_tokens = base + total;
                    }
                }
            }
            bar_sync(bar_a, G);  // #3
        } else if (gt == 0) {
            merged = true;
        }

        // results of the look-back (read before the next barrier; overwritten only after it)
        const unsigned long long rel_base = gs->base;                   // tokens before this tile (this launch)
        const unsigned long long out_base = rel_base + a.out_base_tokens;  // index into a.out
        const uint32_t tile_cin = gs->carry_in;
        const uint32_t f_idx = gs->f_idx, f_delta = gs->f_delta;
        const uint32_t total = gs->total;
        phat_next = gs->carry_out;
        const bool fits = (out_base + total <= a.out_cap_tokens);
        const bool direct = part && fits && (out_base % (SEG / 2) == 0);
        if (!fits && gt == 0) *a.scratch.overflow = 1u;
        // this tile's position bookkeeping is needed below; compute the next tile's first
        TileInfo nti = ti;
        nti.rem0 += stride_elems;
        if (a.chunk != 0 && nti.rem0 >= a.chunk) {
            const unsigned long long q = nti.rem0 / a.chunk;
            nti.ck0 += q;
            nti.rem0 -= q * a.chunk;
        }

        if (direct) {
            // ------------ fast emit: SEG/2 tokens per segment, straight from registers ---------------
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t seg_idx = uint32_t(r * G + gt);
                uint16_t *dst = a.out + out_base + size_t(seg_idx) * (SEG / 2);
                if (SEG == 16) {
                    stg_stream_v4(dst, make_uint4(hv[r][0], hv[r][1], hv[r][HV > 2 ? 2 : 0], hv[r][HV > 3 ? 3 : 0]));
                } else {
                    *reinterpret_cast<uint2 *>(dst) = make_uint2(hv[r][0], hv[r][1]);
                }
            }
            if (a.chunk_ends != nullptr && !(lean && !end_wall)) {  // some segment may hold a chunk end
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t seg_idx = uint32_t(r * G + gt);
                    const uint32_t off = uint32_t(r * ROUND_ELEMS + gt * SEG);
                    const Walls<SEG> wl = seg_walls<SEG, TILE_ELEMS>(a, ti, off, tile_base + off);
                    uint32_t e = wl.endm;
                    unsigned long long ck = wl.ck;
                    while (e) {
                        const int d = __ffs(e) - 1;
                        e &= e - 1;
                        const uint32_t upto = __popc(par_mask & ((2u << d) - 1));
                        a.chunk_ends[ck++] = a.chunk_ends_base + 2ull * (rel_base + seg_idx * (SEG / 2) + upto);
                    }
                }
            }
            ti = nti;
            load_tile(tile + total_groups, w, nx31);  // next tile's input is in flight while the stores drain
            TRACE(6);
            continue;
        }

        // ------------ general emit: compact into the staging buffer ------------------------------------
        const uint32_t phase = uint32_t(out_base & 7);                 // keep the 16-byte phase of the output
#pragma unroll
        for (int r = 0; r < R; ++r) {
            uint32_t em, pos;
            const uint32_t off = uint32_t(r * ROUND_ELEMS + gt * SEG);
            const unsigned long long g = tile_base + off;
            const uint32_t seg_idx = uint32_t(r * G + gt);
            if (part) {  // PART tile whose output is not vector-aligned: same tokens, staged
                em = par_mask;
                pos = seg_idx * (SEG / 2);
                if (phat) {
#pragma unroll
                    for (int k = 0; k < HV; ++k) { ov[r][k] = hv[r][k]; hv[r][k] = 0; }
                } else {
#pragma unroll
                    for (int k = 0; k < HV; ++k) ov[r][k] = 0;
                }
            } else {
                const int q = r * WARPS + wg;
                const uint32_t wexcl = (q < 32) ? __shfl_sync(FULL, wscan0, q & 31) : __shfl_sync(FULL, wscan1, q & 31);
                const uint32_t m = mbits[r] & ALL;
                const bool dep = (mbits[r] >> 31) != 0;
                const uint32_t cin = dep ? tile_cin : ((mbits[r] >> 30) & 1u);
                const uint32_t valid = (g + SEG <= n) ? ALL : (g < n ? ((1u << uint32_t(n - g)) - 1) : 0u);
                const uint32_t st = start_bits(m, cin);
                em = valid & ~((st << 1) | cin);
                pos = wexcl + excl[r] - ((tile_cin && seg_idx > f_idx) ? f_delta : 0u);
            }
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
                    const uint32_t v = (j & 1) ? ov[r][j >> 2] : hv[r][j >> 2];
                    stage[sp++] = uint16_t(((j >> 1) & 1) ? (v >> 16) : v);
                }
            }
        }
        ti = nti;
        load_tile(tile + total_groups, w, nx31);
        bar_sync(bar_b, G);  // #4

        // ------------ staging -> global, 16-byte stores ---------------------------------------------------
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
        }
        // the staging buffer is next written after at least two more barriers of this group
    }
    if (merged) *a.scratch.merged_any = 1u;
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
    // (the dense-abort word at ctrl+384 belongs to the dense pass enqueued in front of this launch: keep it)
    err = cudaMemsetAsync(a.scratch.ctrl, 0, 384, stream);
    if (err != cudaSuccess) return err;
    err = cudaMemsetAsync(a.scratch.tile_status, 0, n_tiles * 8, stream);
    if (err != cudaSuccess) return err;
    const size_t groups = kCtaThreads / G;
    size_t grid = (n_tiles + groups - 1) / groups;
    if (grid > size_t(sms)) grid = size_t(sms);
    if (grid == 0) grid = 1;
    // Cooperative launch: every CTA is resident before any starts, so a group may spin on the
    // descriptors of its (statically assigned) predecessors without risking a deadlock.
    SweepArgs a_arg = a;
    typename FE::Params p_arg = fp;
    void *args[] = {&a_arg, &p_arg};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(kern), dim3(unsigned(grid)), dim3(kCtaThreads),
                                       args, smem, stream);
}

}  // namespace

#ifdef BLT_TRACE
cudaError_t debug_set_trace(unsigned long long *d_buf, unsigned int iters) {
    cudaError_t e = cudaMemcpyToSymbol(g_trace, &d_buf, sizeof d_buf);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(g_trace_iters, &iters, sizeof iters);
}
#endif

// ---- scratch -------------------------------------------------------------------------------------
size_t sweep_scratch_bytes(size_t n_elems_max) {
    const size_t tiles = std::max<size_t>((n_elems_max + kMinTileElems - 1) / kMinTileElems + 1, 2 * 8192);
    return kCtrlBytes + tiles * 8 + tiles * 4;
}
SweepScratch sweep_scratch_carve(void *mem, size_t n_elems_max) {
    SweepScratch s;
    const size_t tiles = std::max<size_t>((n_elems_max + kMinTileElems - 1) / kMinTileElems + 1, 2 * 8192);
    unsigned char *p = static_cast<unsigned char *>(mem);
    s.ctrl = p;
    // The tile counter (one atomic per tile from every SM) and the phase hint (read by every claim)
    // get a 128-byte line each; nothing else may share them, or every tile serialises on that line.
    s.total_tokens = reinterpret_cast<uint64_t *>(p);
    s.merged_any = reinterpret_cast<uint32_t *>(p + 12);
    s.overflow = reinterpret_cast<uint32_t *>(p + 16);
    s.tile_counter = reinterpret_cast<uint32_t *>(p + 128);
    s.phase_hint = reinterpret_cast<uint32_t *>(p + 256);
    s.dense_abort = reinterpret_cast<uint32_t *>(p + 384);
    s.tile_status = reinterpret_cast<uint64_t *>(p + kCtrlBytes);
    s.tile_desc = reinterpret_cast<uint32_t *>(p + kCtrlBytes + tiles * 8);
    s.bytes = kCtrlBytes + tiles * 8 + tiles * 4;
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

static const char *kVariantNames[] = {"g1024r2", "g1024r1", "g512r2", "g256r2", "g512r1", "g256r1"};
int num_sweep_variants() { return int(sizeof(kVariantNames) / sizeof(kVariantNames[0])); }
const char *sweep_variant_name(int v) { return (v >= 0 && v < num_sweep_variants()) ? kVariantNames[v] : "?"; }

cudaError_t launch_bpe_sweep_pairs(const SweepArgs &a_in, const uint16_t *d_table, int variant, cudaStream_t stream) {
    PairsFE::Params p{d_table};
    SweepArgs a = a_in;
    a.dense_flag = nullptr;
    // Dense speculation is sound when no even pair can straddle a wall (even chunk size, or one chunk), the
    // output starts on a 16-byte boundary and the output surely fits.
    static const bool dense_enabled = (getenv("BLT_NO_DENSE") == nullptr);
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
        // the exact launch below zeroes the control block too; the abort word must be clear before the dense pass
        err = cudaMemsetAsync(a.scratch.dense_abort, 0, 4, stream);
        if (err != cudaSuccess) return err;
        const size_t n_segs = a.n / 16;
        size_t grid = (n_segs + kCtaThreads - 1) / kCtaThreads;
        if (grid > size_t(sm_count(dev))) grid = size_t(sm_count(dev));
        if (grid == 0) grid = 1;
        dense_pairs_kernel<<<dim3(unsigned(grid)), dim3(kCtaThreads), PairsFE::TABLE_BYTES, stream>>>(
            static_cast<const unsigned char *>(a.in), a.n, a.out + a.out_base_tokens, d_table, a.scratch.dense_abort);
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
        a.dense_flag = a.scratch.dense_abort;
    }
    static const bool use_lookback = (getenv("BLT_SWEEP_IMPL") != nullptr && std::string(getenv("BLT_SWEEP_IMPL")) == "lookback");
    if (!use_lookback) {
        switch (variant) {
            case 1: return launch_sweep3<PairsFE, 8>(a, p, stream);
            default: return launch_sweep3<PairsFE, 4>(a, p, stream);
        }
    }
    switch (variant) {
        case 1: return launch_sweep<1024, 1, PairsFE>(a, p, stream);
        case 2: return launch_sweep<512, 2, PairsFE>(a, p, stream);
        case 3: return launch_sweep<256, 2, PairsFE>(a, p, stream);
        case 4: return launch_sweep<512, 1, PairsFE>(a, p, stream);
        case 5: return launch_sweep<256, 1, PairsFE>(a, p, stream);
        default: return launch_sweep<1024, 2, PairsFE>(a, p, stream);
    }
}

cudaError_t launch_bpe_sweep_hash(const SweepArgs &a, const HashTableView &t, bool in_is_u16, cudaStream_t stream) {
    if (in_is_u16) {
        HashFE<true>::Params p{t};
        return launch_sweep3<HashFE<true>, 8>(a, p, stream);
    }
    HashFE<false>::Params p{t};
    return launch_sweep3<HashFE<false>, 4>(a, p, stream);
}

}  // namespace bltk
