// detok.cuh -- the inverse of the wire format: big-endian u16 tokens -> bytes.
//
// The reference has no detokenizer; this is the consumer of its output format (tokenizer.rs:88-91) for
// tables with byte keys and distinct ids >= 256 (every merges.txt table): token < 256 -> that byte,
// token = id of (l, r) -> bytes l r.  It is an expansion by 1 or 2 bytes per token, so the output offset
// of a token is a prefix sum of widths.  Same three-launch shape as the exact sweep, no inter-CTA waiting:
//   detok_count_kernel  every warp owns a contiguous range of 512-byte rounds (256 tokens) and counts the
//                       output bytes of its range (tokens + tokens >= 256): a pure streaming read;
//   detok_scan_kernel   one CTA turns the <= 8192 range counts into exclusive offsets and the total;
//   detok_emit_kernel   every warp re-reads its range, looks the ids up in a 128 KiB shared-memory table
//                       (id -> l | r << 8), compacts each round's bytes into a warp-private staging line
//                       and streams whole 16-byte vectors out (the < 16 leftover bytes wait for the next
//                       round; the partial vectors at the two ends of a range go out as byte stores).
// Algorithmic bytes: 2*T_in + N_out (DRAM traffic: 4*T_in + N_out, the token stream is read twice).
// Included by kernels.cu inside its anonymous namespace.
#pragma once

constexpr int kDetokRoundTokens = 256;                       // 32 lanes x 8 tokens (16 bytes per lane)
constexpr int kDetokStageBytes = 2 * kDetokRoundTokens + 16; // per warp
constexpr size_t kDetokSmem = size_t(kPairTableEntries) * 2 + 8192 + size_t(kCtaThreads / 32) * kDetokStageBytes;

struct DetokWalk {
    long long cur, end;
    __device__ __forceinline__ void init(size_t n_tok, long long warp, long long n_warps) {
        const long long n_rounds = (long long)((n_tok + kDetokRoundTokens - 1) / kDetokRoundTokens);
        const long long per = (n_rounds + n_warps - 1) / n_warps;
        cur = warp * per;
        end = cur + per < n_rounds ? cur + per : n_rounds;
    }
};

// The lane's 8 tokens of a round (as stored: big-endian u16, i.e. the token's HIGH byte is the LOW byte of
// each little-endian half-word); missing tokens of a ragged last round read as zero.  *n_valid = how many exist.
__device__ __forceinline__ uint4 detok_load(const DetokArgs &a, long long round, int lane, uint32_t *n_valid) {
    const unsigned long long t0 = (unsigned long long)round * kDetokRoundTokens + uint32_t(lane) * 8u;
    if (t0 + 8 <= a.n_tok) {
        *n_valid = 8;
        return ldg_stream_v4(a.in + t0);
    }
    uint32_t tmp[4] = {0, 0, 0, 0};
    uint32_t k = 0;
    for (; k < 8 && t0 + k < a.n_tok; ++k) tmp[k >> 1] |= uint32_t(a.in[t0 + k]) << (16 * (k & 1));
    *n_valid = k;
    return make_uint4(tmp[0], tmp[1], tmp[2], tmp[3]);
}

// Which of the lane's 8 tokens are >= 256 (the token's high byte, the low byte of the stored half-word, is non-zero):
// bit 7 of byte j of f[0] <-> token j, of f[1] <-> token 4 + j (a SWAR "byte != 0"; the flags are used where they are)
struct DetokWide {
    uint32_t f[2];
    __device__ __forceinline__ uint32_t count() const { return __popc(f[0]) + __popc(f[1]); }
};
__device__ __forceinline__ uint32_t detok_nonzero_flags(uint32_t x) { return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }
__device__ __forceinline__ DetokWide detok_wide(const uint4 &w) {
    DetokWide d;
    d.f[0] = detok_nonzero_flags(__byte_perm(w.x, w.y, 0x6420));
    d.f[1] = detok_nonzero_flags(__byte_perm(w.z, w.w, 0x6420));
    return d;
}

// R consecutive rounds starting at round r0: R independent 16-byte loads per lane.  Rounds from r_end on read as
// absent (nv = 0); the stream's ragged last round goes through detok_load.
template <int R>
__device__ __forceinline__ void detok_load_range(const DetokArgs &a, long long r0, long long r_end, int lane, uint4 *wq, uint32_t *nvq) {
    if (r0 + R <= r_end && (unsigned long long)(r0 + R) * kDetokRoundTokens <= a.n_tok) {  // warp-uniform: every round is whole
        const uint16_t *p = a.in + (unsigned long long)r0 * kDetokRoundTokens + uint32_t(lane) * 8u;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            wq[r] = ldg_stream_v4(p + r * kDetokRoundTokens);
            nvq[r] = 8;
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        nvq[r] = 0;
        wq[r] = make_uint4(0, 0, 0, 0);
        if (r0 + r < r_end) wq[r] = detok_load(a, r0 + r, lane, &nvq[r]);
    }
}

__global__ void __launch_bounds__(kCtaThreads, 1) detok_count_kernel(const DetokArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (kCtaThreads / 32) + (threadIdx.x >> 5);
    DetokWalk wk;
    wk.init(a.n_tok, warp, (long long)gridDim.x * (kCtaThreads / 32));
    unsigned long long bytes = 0;
    constexpr int RB = 8;  // independent 16-byte loads in flight per lane
    for (; wk.cur < wk.end; wk.cur += RB) {
        uint4 wq[RB];
        uint32_t nvq[RB];
        detok_load_range<RB>(a, wk.cur, wk.end, lane, wq, nvq);
#pragma unroll
        for (int r = 0; r < RB; ++r) bytes += nvq[r] + detok_wide(wq[r]).count();  // absent tokens are zero: never wide
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) bytes += __shfl_xor_sync(FULL, bytes, d);
    if (lane == 0 && warp < kMaxRanges) a.scratch.tile_status[kMaxRanges + warp] = bytes;
}

__global__ void __launch_bounds__(kCtaThreads, 1) detok_scan_kernel(const DetokArgs a, int n_ranges) {
    __shared__ unsigned long long warp_sum[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int t0 = threadIdx.x * kScanItems;
    unsigned long long item[kScanItems], agg = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        item[i] = (t0 + i < n_ranges) ? a.scratch.tile_status[kMaxRanges + t0 + i] : 0ull;
        agg += item[i];
    }
    unsigned long long inc = agg;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const unsigned long long o = __shfl_up_sync(FULL, inc, s);
        if (lane >= s) inc += o;
    }
    if (lane == 31) warp_sum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned long long v = warp_sum[lane], x = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const unsigned long long o = __shfl_up_sync(FULL, x, s);
            if (lane >= s) x += o;
        }
        warp_sum[lane] = x - v;  // exclusive
    }
    __syncthreads();
    unsigned long long b = warp_sum[wid] + inc - agg;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (t0 + i < n_ranges) a.scratch.tile_status[t0 + i] = b;
        b += item[i];
    }
    if (threadIdx.x == kCtaThreads - 1) {
        *a.scratch.total_tokens = b;  // output BYTES for this launch
        if (b > a.out_cap) *a.scratch.overflow = 1u;
    }
}

// One warp's writer: turns rounds of 256 tokens (8 per lane) into bytes at a given output offset.  The bytes of a
// round are compacted into the warp's staging line and whole 16-byte vectors are streamed out; the < 16 leftover bytes
// wait for the next round.  The partial vectors at the two ends of the warp's range (shared with the neighbouring
// ranges) go out as byte stores.  Used by the three-launch emit kernel (one long range per warp) and by the fused
// kernel (one short range per warp and tile).
// Table: dec[id] = l | r << 8 for a merged id, dec[b] = b for b < 256, so a token's bytes are dec[id] whatever its
// width; the width is 1 + (id >= 256).  Ids >= limit are caught by a packed running maximum (checked in finish()),
// holes below the limit by the bitmap (HOLES only; the bits of ids < 256 are set).
template <bool HOLES>
struct DetokEmitter {
    const DetokArgs &a;
    const unsigned char *tbl;   // shared: the table
    const uint32_t *exists;     // shared: the bitmap (HOLES)
    unsigned char *stage;       // shared: this warp's staging line
    const int lane;
    unsigned char *op = nullptr;  // stage[0] is *op, a 16-byte boundary of the output
    uint32_t pend = 0;            // stage[0 .. pend) holds bytes not yet written
    uint32_t head = 0;            // the first `head` bytes of the range's first vector belong to the range in front
    bool bad = false;
    uint32_t idmax = 0;           // packed running maximum of the ids seen

    __device__ __forceinline__ DetokEmitter(const DetokArgs &a_, const unsigned char *tbl_, const uint32_t *ex_, unsigned char *stage_, int lane_)
        : a(a_), tbl(tbl_), exists(ex_), stage(stage_), lane(lane_) {}

    __device__ __forceinline__ void begin(unsigned long long base) {
        op = a.out + (base & ~15ull);
        pend = uint32_t(base & 15ull);
        head = pend;
    }
    __device__ __forceinline__ uint32_t dec16(uint32_t byte_off) const {
        return *reinterpret_cast<const uint16_t *>(tbl + byte_off);
    }
    __device__ __forceinline__ bool hole(uint32_t id) const { return ((exists[id >> 5] >> (id & 31)) & 1u) == 0u; }

    // One round: the lane's 8 tokens as loaded (big-endian halves), nv of them exist (8 in every round but the stream's last).
    __device__ __forceinline__ void round(const uint4 &w, uint32_t nv) {
        const DetokWide wide = detok_wide(w);
        const uint32_t cnt = nv + wide.count();
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        const uint32_t words[4] = {w.x, w.y, w.z, w.w};
        if (__all_sync(FULL, nv == 8)) {
            // ---- full round: 8 lookups; the pair (2k, 2k+1) is packed with one byte permute chosen by the width of
            // token 2k, the four pairs with funnel shifts by their widths ----
            uint32_t c[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t ids = __byte_perm(words[k], 0, 0x2301);  // both halves byte-swapped: the two ids
                idmax = __vmaxu2(idmax, ids);
                const uint32_t e0 = dec16((ids << 1) & 0x1FFFEu), e1 = dec16((ids >> 15) & 0x1FFFEu);
                if (HOLES) bad = bad || hole(ids & 0xffffu) || hole(ids >> 16);
                c[k] = __byte_perm(e0, e1, (wide.f[k >> 1] & (0x80u << (16 * (k & 1)))) ? 0x5410u : 0x3540u);  // e0's bytes 2, 3 are zero
            }
            if (pend == 0 && total == 512u) {  // every token is a merged id and the output is vector-aligned: straight out
                stg_stream_v4(op + 16 * lane, make_uint4(c[0], c[1], c[2], c[3]));
                op += 512;
                return;
            }
            const uint32_t w01 = 16u + 8u * __popc(wide.f[0] & 0x8080u), w45 = 16u + 8u * __popc(wide.f[1] & 0x8080u);  // 16..32 bits
            // d0 = c0 | c1 << w01, d1 = c2 | c3 << w45 (64-bit each, as two words; clamped funnel shifts)
            const uint32_t d0l = c[0] | __funnelshift_lc(0u, c[1], w01), d0h = __funnelshift_lc(c[1], 0u, w01);
            const uint32_t d1l = c[2] | __funnelshift_lc(0u, c[3], w45), d1h = __funnelshift_lc(c[3], 0u, w45);
            const uint32_t sh2 = 8u * __popc(wide.f[0]);  // S = d0 | d1 << (32 + sh2), sh2 in 0..32
            const uint32_t s0 = d0l, s1 = d0h | __funnelshift_lc(0u, d1l, sh2);
            const uint32_t s2 = __funnelshift_lc(d1l, d1h, sh2), s3 = __funnelshift_lc(d1h, 0u, sh2);
            if (pend == 0 && total == 256u) {  // every token is a plain byte and the output is vector-aligned: straight out
                *reinterpret_cast<uint2 *>(op + 8 * lane) = make_uint2(s0, s1);
                op += 256;
                return;
            }
            // place the lane's cnt bytes at byte offset o of the staging line with 4-byte stores: shift by o & 3,
            // take the previous lane's incomplete last word into my first one, store my complete words
            const uint32_t o = pend + (incl - cnt);
            const uint32_t sh = (o & 3u) * 8u;
            uint32_t t0 = s0 << sh;
            const uint32_t t1 = __funnelshift_l(s0, s1, sh), t2 = __funnelshift_l(s1, s2, sh);
            const uint32_t t3 = __funnelshift_l(s2, s3, sh), t4 = __funnelshift_l(s3, 0u, sh);
            const uint32_t cw = ((o + cnt) >> 2) - (o >> 2);  // complete words: 2..4
            const bool ragged_end = ((o + cnt) & 3u) != 0u;
            const uint32_t my_tail = !ragged_end ? 0u : (cw == 2 ? t2 : cw == 3 ? t3 : t4);
            uint32_t prev = __shfl_up_sync(FULL, my_tail, 1);
            uint32_t *sw = reinterpret_cast<uint32_t *>(stage) + (o >> 2);
            const uint32_t left = sw[0] & ~(0xffffffffu << sh);  // lane 0: the bytes the previous round left in its first word
            t0 |= (lane == 0) ? left : prev;
            sw[0] = t0;
            sw[1] = t1;
            if (cw > 2) sw[2] = t2;
            if (cw > 3) sw[3] = t3;
            if (lane == 31 && ragged_end) sw[cw] = my_tail;
        } else {
            // ---- the ragged last round of the stream: byte stores ----
            unsigned char *sp = stage + pend + (incl - cnt);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (uint32_t(k) < nv) {
                    const uint32_t half = (words[k >> 1] >> (16 * (k & 1))) & 0xffffu;
                    const uint32_t id = __byte_perm(half, 0, 0x4401);
                    bad = bad || id >= a.limit || (HOLES && hole(id));
                    const uint32_t e = dec16(id << 1);
                    sp[0] = static_cast<unsigned char>(e);
                    if (wide.f[k >> 2] & (0x80u << (8 * (k & 3)))) {
                        sp[1] = static_cast<unsigned char>(e >> 8);
                        sp += 2;
                    } else {
                        sp += 1;
                    }
                }
            }
        }
        __syncwarp();
        // flush the whole 16-byte vectors, keep the leftover (< 16 bytes) at the front of the line
        const uint32_t have = pend + total;
        const uint32_t nvec = have >> 4;
        if (nvec != 0 && head != 0) {  // the range's first vector: only the bytes from `head` on are this range's
            if (uint32_t(lane) >= head && lane < 16) op[lane] = stage[lane];
        }
        // (have <= 15 + 512: at most 32 whole vectors, one per lane)
        if (uint32_t(lane) < nvec && (lane != 0 || head == 0)) stg_stream_v4(op + 16 * lane, *reinterpret_cast<const uint4 *>(stage + 16 * lane));
        const uint32_t rem = have & 15u;
        unsigned char keep = 0;
        if (nvec != 0 && uint32_t(lane) < rem) keep = stage[16 * nvec + lane];
        __syncwarp();
        if (nvec != 0) {
            if (uint32_t(lane) < rem) stage[lane] = keep;
            head = 0;
            op += 16u * nvec;
            pend = rem;
        } else {
            pend = have;
        }
        __syncwarp();
    }

    // the tail of the range: bytes [head, pend) of a vector shared with the next range; true = an unknown token was seen
    __device__ __forceinline__ bool finish() {
        if (head + uint32_t(lane) < pend) op[head + lane] = stage[head + lane];  // pend < 16 at the end of a round
        bad = bad || (idmax & 0xffffu) >= a.limit || (idmax >> 16) >= a.limit;
        return __any_sync(FULL, bad);
    }
};

__device__ __forceinline__ void detok_load_table(const DetokArgs &a, unsigned char *smem, bool holes) {
    const uint4 *src = reinterpret_cast<const uint4 *>(a.table);
    uint4 *dst = reinterpret_cast<uint4 *>(smem);
    const int n16 = (kPairTableEntries * 2 + (holes ? 8192 : 0)) / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
}

template <bool HOLES>
__global__ void __launch_bounds__(kCtaThreads, 1) detok_emit_kernel(const DetokArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    detok_load_table(a, smem, HOLES);
    __syncthreads();
    if (*reinterpret_cast<const volatile uint32_t *>(a.scratch.overflow) != 0u) return;  // the scan found it does not fit
    const int lane = threadIdx.x & 31;
    unsigned char *stage = smem + kPairTableEntries * 2 + 8192 + size_t(threadIdx.x >> 5) * kDetokStageBytes;
    const long long warp = (long long)blockIdx.x * (kCtaThreads / 32) + (threadIdx.x >> 5);
    DetokWalk wk;
    wk.init(a.n_tok, warp, (long long)gridDim.x * (kCtaThreads / 32));
    if (wk.cur >= wk.end || warp >= kMaxRanges) return;
    DetokEmitter<HOLES> em(a, smem, reinterpret_cast<const uint32_t *>(smem + kPairTableEntries * 2), stage, lane);
    em.begin(a.scratch.tile_status[warp]);
    constexpr int RB = 4;  // rounds loaded together: four independent 16-byte loads in flight per lane
    for (; wk.cur < wk.end; wk.cur += RB) {
        uint4 wq[RB];
        uint32_t nvq[RB];
        detok_load_range<RB>(a, wk.cur, wk.end, lane, wq, nvq);
#pragma unroll
        for (int r = 0; r < RB; ++r)
            if (wk.cur + r < wk.end) em.round(wq[r], nvq[r]);
    }
    if (em.finish() && lane == 0) reinterpret_cast<uint32_t *>(a.scratch.ctrl)[6] = 1u;  // "unknown token"
}

// ---- the fused form: count, decoupled look-back and emit in ONE launch -----------------------------------------
// One persistent CTA per SM (the table takes 128 of its 227 KiB), split into two GROUPS of 16 warps that work on
// different tiles and meet on their own named barrier, so that one group's waits (DRAM latency of the count, the
// barrier, the look-back) fall into the other's emit phase.  A tile is 16 warps x kDfRounds rounds = 16 384 tokens,
// dealt round-robin over the 2 x grid groups.  Iteration i of a group: its warp 0 resolves the look-back of tile t_i
// (whose byte count it published one iteration ago, so its predecessors have had a whole emit phase to publish theirs)
// while all its warps count tile t_i+1 (a streaming read that leaves the tile in L2); one barrier; warp 0 publishes
// t_i+1's count at once; then every warp emits its range of t_i from the tokens it re-read from L2 in front of the
// count.  DRAM traffic = 2*T_in + N_out, the algorithmic bytes.  A tile only waits for tiles with smaller numbers, which
// belong to resident CTAs (grid <= SM count; launches that wait on other CTAs are chained per device, see
// FusedLaunch::launch) and nothing that publishes a count waits.
constexpr int kDfRounds = 4;                                          // rounds per warp and tile
#ifndef BLT_DF_GROUPS
#define BLT_DF_GROUPS 1
#endif
constexpr int kDfGroups = BLT_DF_GROUPS;                              // groups per CTA (2 measured slower: smaller tiles, longer window)
constexpr int kDfAhead = 1;                                           // a tile is counted this many iterations before it is emitted (2 measured slower: L2 misses)
constexpr int kDfGroupWarps = (kCtaThreads / 32) / kDfGroups;         // 16
constexpr int kDfTileRounds = kDfGroupWarps * kDfRounds;              // 64 rounds = 32 KiB of tokens
constexpr int kDfLB = kDfGroups == 1 ? 6 : 10;                        // look-back window: 32 x this many tiles (> one round of the deal)
constexpr unsigned long long DF_A = 1ull << 62, DF_P = 2ull << 62, DF_VAL = (1ull << 62) - 1;

struct DetokFusedShared {             // one per group
    // slot = iteration % 4
    unsigned long long wexcl[4][kDfGroupWarps];  // bytes of the tile in front of each warp's range
    unsigned long long base[4];                  // bytes of the stream in front of the tile
    uint32_t wtot[4][kDfGroupWarps];             // bytes of each warp's range
    uint32_t skip[4];                            // the tile does not fit into the output
};
constexpr size_t kDetokFusedSmem = kDetokSmem + 2048;
static_assert(kDfGroups * sizeof(DetokFusedShared) <= 2048, "group state");

template <bool HOLES>
__global__ void __launch_bounds__(kCtaThreads, 1)
detok_fused_kernel(const DetokArgs a, unsigned long long *__restrict__ desc, uint32_t n_tiles) {
    extern __shared__ __align__(16) unsigned char smem[];
    detok_load_table(a, smem, HOLES);
    const int lane = threadIdx.x & 31;
    const int grp = (threadIdx.x >> 5) / kDfGroupWarps, wid = (threadIdx.x >> 5) % kDfGroupWarps;  // group, warp in the group
    unsigned char *stage = smem + kPairTableEntries * 2 + 8192 + size_t(threadIdx.x >> 5) * kDetokStageBytes;
    DetokFusedShared *sh = reinterpret_cast<DetokFusedShared *>(smem + kDetokSmem) + grp;
    const long long n_rounds = (long long)((a.n_tok + kDetokRoundTokens - 1) / kDetokRoundTokens);
    auto group_barrier = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(kDfGroupWarps * 32) : "memory"); };

    // every warp counts the bytes of its range of tile t (buffer b)
    auto count_tile = [&](uint32_t t, int b) {
        const long long r0 = (long long)t * kDfTileRounds + (long long)wid * kDfRounds;
        uint4 wq[kDfRounds];
        uint32_t nvq[kDfRounds];
        detok_load_range<kDfRounds>(a, r0, n_rounds, lane, wq, nvq);
        uint32_t bytes = 0;
#pragma unroll
        for (int r = 0; r < kDfRounds; ++r) bytes += nvq[r] + detok_wide(wq[r]).count();  // absent tokens are zero: never wide
        bytes = __reduce_add_sync(FULL, bytes);
        if (lane == 0) sh->wtot[b][wid] = bytes;
    };
    // warp 0: the ranges' offsets inside the tile, and the tile's count published (status A); returns the count
    auto publish_count = [&](uint32_t t, int b) -> unsigned long long {
        const uint32_t v = lane < kDfGroupWarps ? sh->wtot[b][lane] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < kDfGroupWarps; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane < kDfGroupWarps) sh->wexcl[b][lane] = inc - v;
        const unsigned long long agg = __shfl_sync(FULL, inc, kDfGroupWarps - 1);
        if (lane == 0) st_desc(desc + t, DF_A | agg);
        return agg;
    };
    // warp 0: bytes of the stream in front of tile t; publishes the tile's inclusive prefix (status P)
    auto look_back = [&](uint32_t t, unsigned long long agg, int b) {
        unsigned long long excl = 0;
        bool ok = false;
        for (uint32_t tries = 0; tries < (1u << 22) && !ok; ++tries) {
            unsigned long long d[kDfLB];
#pragma unroll
            for (int j = 0; j < kDfLB; ++j) {
                const long long idx = (long long)t - 1 - lane - 32 * j;
                d[j] = DF_P;  // in front of tile 0: nothing
                if (idx >= 0) d[j] = ld_desc(desc + idx);
            }
            // the nearest inclusive prefix with nothing missing in front of it (every branch is warp-uniform)
            unsigned long long sum = 0;
            bool hole = false, found = false;
#pragma unroll
            for (int j = 0; j < kDfLB; ++j) {
                if (!found && !hole) {
                    const uint32_t st = uint32_t(d[j] >> 62);
                    const uint32_t pm = __ballot_sync(FULL, st == 2u), zm = __ballot_sync(FULL, st == 0u);
                    const uint32_t q = pm ? uint32_t(__ffs(pm) - 1) : 32u;           // lanes < q are nearer than the prefix
                    const uint32_t nearer = q >= 32u ? FULL : ((1u << q) - 1u);
                    if (zm & nearer) { hole = true; }
                    else {
                        const bool take = (uint32_t(lane) < q) || (uint32_t(lane) == q);
                        unsigned long long part = take ? (d[j] & DF_VAL) : 0ull;
#pragma unroll
                        for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(FULL, part, s);
                        sum += part;
                        if (pm) found = true;
                    }
                }
            }
            if (found) { excl = sum; ok = true; }
            else __nanosleep(40);
        }
        const unsigned long long incl = excl + agg;
        if (lane == 0) {
            uint32_t skip = 0;
            if (!ok) { *a.scratch.overflow = 3u; skip = 1; }
            if (incl > a.out_cap) { *a.scratch.overflow = 1u; skip = 1; }
            st_desc(desc + t, DF_P | incl);
            sh->base[b] = excl;
            sh->skip[b] = skip;
            if (t == n_tiles - 1) *a.scratch.total_tokens = incl;  // output BYTES of this launch
        }
    };

    const uint32_t step = gridDim.x * kDfGroups;  // tiles are dealt to the groups of all CTAs round-robin
    uint32_t t = blockIdx.x * kDfGroups + grp;
    // tiles are counted (and their counts published) kDfAhead iterations before they are emitted
    unsigned long long agg[kDfAhead];
#pragma unroll
    for (int d = 0; d < kDfAhead; ++d) {
        agg[d] = 0;
        const unsigned long long td = (unsigned long long)t + (unsigned long long)d * step;
        if (td < n_tiles) count_tile(uint32_t(td), d);
    }
    __syncthreads();  // the table and the first counts
    if (wid == 0) {
#pragma unroll
        for (int d = 0; d < kDfAhead; ++d) {
            const unsigned long long td = (unsigned long long)t + (unsigned long long)d * step;
            if (td < n_tiles) agg[d] = publish_count(uint32_t(td), d);
        }
    }
    DetokEmitter<HOLES> em(a, smem, reinterpret_cast<const uint32_t *>(smem + kPairTableEntries * 2), stage, lane);
    bool bad = false;
    for (uint32_t i = 0; t < n_tiles; ++i) {
        const int cb = int(i & 3u), nb = int((i + kDfAhead) & 3u);
        const unsigned long long t_next = (unsigned long long)t + (unsigned long long)kDfAhead * step;
        const bool have_next = t_next < n_tiles;
        // the tokens of this warp's range of tile t, again (from L2: the tile was counted kDfAhead iterations ago); the
        // loads are issued in front of the count so that they have landed when the barrier opens
        const long long r0 = (long long)t * kDfTileRounds + (long long)wid * kDfRounds;
        uint4 wq[kDfRounds];
        uint32_t nvq[kDfRounds];
        if (r0 < n_rounds) detok_load_range<kDfRounds>(a, r0, n_rounds, lane, wq, nvq);
        if (wid == 0) look_back(t, agg[0], cb);
        if (have_next) count_tile(uint32_t(t_next), nb);
        group_barrier();
        if (wid == 0) {
#pragma unroll
            for (int d = 0; d + 1 < kDfAhead; ++d) agg[d] = agg[d + 1];
            if (have_next) agg[kDfAhead - 1] = publish_count(uint32_t(t_next), nb);
        }
        if (sh->skip[cb] == 0u && r0 < n_rounds) {
            em.begin(sh->base[cb] + sh->wexcl[cb][wid]);
#pragma unroll
            for (int r = 0; r < kDfRounds; ++r)
                if (r0 + r < n_rounds) em.round(wq[r], nvq[r]);
            bad = em.finish() || bad;
        }
        if ((unsigned long long)t + step >= n_tiles) break;
        t += step;
    }
    if (bad && lane == 0) reinterpret_cast<uint32_t *>(a.scratch.ctrl)[6] = 1u;  // "unknown token"
}

// variant: 0 = three launches (count, scan, emit), 1 = the fused single launch; *launches = kernels launched
cudaError_t launch_detok_impl(const DetokArgs &a, int variant, cudaStream_t stream, int *launches) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev >= kMaxDevices) return cudaErrorInvalidDevice;
    static std::atomic<bool> configured[kMaxDevices];
    if (!configured[dev].load(std::memory_order_acquire)) {
        err = cudaFuncSetAttribute(detok_emit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kDetokSmem));
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(detok_emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kDetokSmem));
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(detok_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kDetokFusedSmem));
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(detok_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kDetokFusedSmem));
        if (err != cudaSuccess) return err;
        configured[dev].store(true, std::memory_order_release);
    }
    err = cudaMemsetAsync(a.scratch.ctrl, 0, 256, stream);
    if (err != cudaSuccess) return err;
    const size_t n_rounds = (a.n_tok + kDetokRoundTokens - 1) / kDetokRoundTokens;
    const size_t n_tiles = (n_rounds + kDfTileRounds - 1) / kDfTileRounds;
    if (variant != 0 && n_tiles * 8 <= a.scratch.meta_bytes && n_tiles < 0xfffffff0ull) {
        err = cudaMemsetAsync(a.scratch.meta, 0, n_tiles * 8, stream);
        if (err != cudaSuccess) return err;
        const size_t grid = std::min((n_tiles + kDfGroups - 1) / kDfGroups, size_t(sm_count(dev)));
        // CTAs wait for tiles of other CTAs of the launch: chained per device like the fused sweep (FusedLaunch::launch)
        FusedChain &fc = fused_chain(dev);
        std::lock_guard<std::mutex> lk(fc.mu);
        if (fc.ev == nullptr) {
            err = cudaEventCreateWithFlags(&fc.ev, cudaEventDisableTiming);
            if (err != cudaSuccess) return err;
        } else {
            err = cudaStreamWaitEvent(stream, fc.ev, 0);
            if (err != cudaSuccess) return err;
        }
        unsigned long long *desc = reinterpret_cast<unsigned long long *>(a.scratch.meta);
        if (a.holes) detok_fused_kernel<true><<<dim3(unsigned(grid)), dim3(kCtaThreads), kDetokFusedSmem, stream>>>(a, desc, uint32_t(n_tiles));
        else detok_fused_kernel<false><<<dim3(unsigned(grid)), dim3(kCtaThreads), kDetokFusedSmem, stream>>>(a, desc, uint32_t(n_tiles));
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
        if (launches) *launches = 1;
        return cudaEventRecord(fc.ev, stream);
    }
    if (a.scratch.max_tiles < size_t(2 * kMaxRanges)) return cudaErrorInvalidValue;
    const size_t warps_per_cta = kCtaThreads / 32;
    size_t grid = (n_rounds + warps_per_cta - 1) / warps_per_cta;
    if (grid > size_t(sm_count(dev))) grid = size_t(sm_count(dev));
    if (grid > size_t(kMaxRanges) / warps_per_cta) grid = size_t(kMaxRanges) / warps_per_cta;
    if (grid == 0) grid = 1;
    detok_count_kernel<<<dim3(unsigned(grid)), dim3(kCtaThreads), 0, stream>>>(a);
    detok_scan_kernel<<<1, kCtaThreads, 0, stream>>>(a, int(grid * warps_per_cta));
    if (a.holes) detok_emit_kernel<true><<<dim3(unsigned(grid)), dim3(kCtaThreads), kDetokSmem, stream>>>(a);
    else detok_emit_kernel<false><<<dim3(unsigned(grid)), dim3(kCtaThreads), kDetokSmem, stream>>>(a);
    if (launches) *launches = 3;
    return cudaGetLastError();
}
