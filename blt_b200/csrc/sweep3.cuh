// sweep3.cuh -- the exact sweep as three launches without any inter-CTA waiting:
//
//   count_kernel  every warp owns a CONTIGUOUS range of tiles (a tile = R rounds x 32 lanes x 16 bytes),
//                 looks all pairs up and reduces its whole range to one carry function: identity or
//                 constant carry_out, tokens emitted if carry_in = 0, and the 0/1-token `delta` for
//                 carry_in = 1.
//   scan_kernel   one CTA composes the (at most 8192) range functions in order and stores, per warp,
//                 the carry entering its range and the number of tokens emitted before it (chunk walls
//                 need nothing special: a wall makes a function constant).
//   emit_kernel   every warp re-reads its range knowing carry_in and base and streams its output.  It
//                 looks up the pairs of the carry's parity first; where all 32 lanes find theirs the
//                 warp-round is dense and its tokens go out with one 16-byte store per lane straight
//                 from registers; otherwise the other parity is looked up and the round is compacted
//                 into a warp-private shared-memory staging line that only ever flushes whole 16-byte
//                 vectors (the < 8 leftover tokens stay for the next round).
//
// Everything is warp-synchronous (no block barrier after the table load, no spinning, no cooperative
// launch); the price is that the input is read twice (2*N_in + 2*T_out bytes of DRAM traffic).
// Included by kernels.cu inside its anonymous namespace, after the front ends and seg_walls().
#pragma once

constexpr uint32_t D_ID = 1u << 31, D_CONST = 1u << 30, D_DELTA = 1u << 29, D_CNT = (1u << 24) - 1;
constexpr uint64_t R_CARRY = 1ull << 63;

template <class FE, int R>
struct Sweep3Cfg {
    static constexpr int SEG = FE::SEG;
    static constexpr int HV = SEG / 4;
    static constexpr int ROUND_ELEMS = 32 * SEG;
    static constexpr int TILE_ELEMS = R * ROUND_ELEMS;
    static constexpr uint32_t ALL = (1u << SEG) - 1;
    static constexpr uint32_t EVEN = 0x55555555u & ALL;
    static constexpr uint32_t HALF_ALL = (1u << (SEG / 2)) - 1;
    static constexpr int STAGE_TOKENS = ROUND_ELEMS + 8;  // per warp
};

// Per-warp walk over its contiguous range of tiles, with the chunk bookkeeping kept incrementally.
template <int TILE_ELEMS>
struct TileWalk {
    long long tile, end;
    TileInfo ti;
    __device__ __forceinline__ void init(const SweepArgs &a, long long warp, long long n_warps) {
        const long long n_tiles = (long long)((a.n + TILE_ELEMS - 1) / TILE_ELEMS);
        const long long per = (n_tiles + n_warps - 1) / n_warps;
        tile = warp * per;
        end = tile + per < n_tiles ? tile + per : n_tiles;
        ti.rem0 = (unsigned long long)tile * TILE_ELEMS;
        ti.ck0 = 0;
        if (a.chunk != 0 && tile < end) { ti.ck0 = ti.rem0 / a.chunk; ti.rem0 -= ti.ck0 * a.chunk; }
    }
    __device__ __forceinline__ void next(const SweepArgs &a) {
        tile += 1;
        ti.rem0 += TILE_ELEMS;
        if (a.chunk != 0 && ti.rem0 >= a.chunk) {
            const unsigned long long q = ti.rem0 / a.chunk;
            ti.ck0 += q;
            ti.rem0 -= q * a.chunk;
        }
    }
};

// Loads one tile's R segments for this lane (+ the look-ahead element of every segment in lane 31).
template <class FE, int R>
__device__ __forceinline__ void load_tile3(const SweepArgs &a, unsigned long long tile_base, int lane, uint4 *w, uint32_t *nx) {
    using C = Sweep3Cfg<FE, R>;
    if (tile_base + C::TILE_ELEMS + 1 <= a.n) {
        // interior tile (all but the last one or two of a launch): one 64-bit address, constant offsets,
        // no bounds checks; the look-ahead elements exist
        const unsigned char *p = static_cast<const unsigned char *>(a.in) + (tile_base + uint32_t(lane * C::SEG)) * FE::ELEM;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            w[r] = ldg_stream_v4(p + r * 512);
            nx[r] = 0;
        }
        if (lane == 31) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (FE::ELEM == 1) nx[r] = p[r * 512 + 16];
                else nx[r] = __byte_perm(uint32_t(*reinterpret_cast<const uint16_t *>(p + r * 512 + 16)), 0, 0x4401);
            }
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const unsigned long long g = tile_base + uint32_t(r * C::ROUND_ELEMS + lane * C::SEG);
        w[r] = make_uint4(0, 0, 0, 0);
        nx[r] = 0;
        if (g + C::SEG <= a.n) {
            w[r] = ldg_stream_v4(static_cast<const unsigned char *>(a.in) + g * FE::ELEM);
        } else if (g < a.n) {  // ragged last segment: element-wise, never reads past n
            uint32_t tmp[4] = {0, 0, 0, 0};
            for (int j = 0; j < C::SEG && g + j < a.n; ++j) {
                const uint32_t v = FE::load_elem(a.in, g + j);
                if (FE::ELEM == 1) tmp[j >> 2] |= v << (8 * (j & 3));
                else tmp[j >> 1] |= __byte_perm(v, 0, 0x4401) << (16 * (j & 1));
            }
            w[r] = make_uint4(tmp[0], tmp[1], tmp[2], tmp[3]);
        }
        if (lane == 31 && g + C::SEG < a.n) nx[r] = FE::load_elem(a.in, g + C::SEG);
    }
}

// Membership word of one segment (both parities looked up), with walls and the end of input applied.
// hv / ov receive the tokens of the even / odd positions (raw token at wall positions).
// have_par >= 0: the tokens of that parity are already in hv (from a dense attempt) and are not looked up again.
template <class FE, int TILE_ELEMS>
__device__ __forceinline__ uint32_t segment_full(const FE &fe, const SweepArgs &a, const TileInfo &ti, bool simple,
                                                 bool end_wall_here, uint32_t off, unsigned long long g, const uint4 &w,
                                                 uint32_t next, uint32_t *hv, uint32_t *ov, uint32_t *valid_out,
                                                 Walls<FE::SEG> *wl_out, int have_par = -1) {
    constexpr int SEG = FE::SEG;
    constexpr uint32_t ALL = (1u << SEG) - 1;
    uint32_t hp, op;
    if (FE::kMembershipInValue && have_par >= 0) {
        if (have_par == 1) {  // hv holds the odd positions: move them over and look the even ones up
#pragma unroll
            for (int k = 0; k < SEG / 4; ++k) ov[k] = hv[k];
            fe.lookup_vals(w, next, 0u, hv);
        } else {
            fe.lookup_vals(w, next, 1u, ov);
        }
        hp = FE::present_mask(hv);
        op = FE::present_mask(ov);
    } else {
        hp = fe.lookup_half(w, next, 0u, hv);
        op = fe.lookup_half(w, next, 1u, ov);
    }
    Walls<SEG> wl;
    wl.endm = 0;
    wl.ck = ti.ck0;
    uint32_t valid = ALL;
    if (simple) {
        if (end_wall_here) wl.endm = 1u << (SEG - 1);
    } else {
        wl = seg_walls<SEG, TILE_ELEMS>(a, ti, off, g);
        valid = (g + SEG <= a.n) ? ALL : (g < a.n ? ((1u << uint32_t(a.n - g)) - 1) : 0u);
        wl.endm &= valid;  // chunk walls behind the end of the input do not exist (they would index past chunk_ends)
    }
    if (wl.endm) {  // a wall suppresses the pair: the raw token is emitted there, not the merged id
#pragma unroll
        for (int j = 0; j < SEG; ++j) {
            if ((wl.endm >> j) & 1u) {
                const uint32_t be = FE::raw_be(w, j);
                uint32_t &dst = (j & 1) ? ov[j >> 2] : hv[j >> 2];
                dst = ((j >> 1) & 1) ? ((dst & 0x0000ffffu) | (be << 16)) : ((dst & 0xffff0000u) | be);
            }
        }
    }
    *valid_out = valid;
    *wl_out = wl;
    return (spread_even(hp) | (spread_even(op) << 1)) & valid & ~wl.endm & ALL;
}

// ---------------------------------------------------------------------------------------------------
constexpr int kMaxRanges = 8192;  // warps of one launch (148 SMs x 32 warps = 4736)

template <class FE, int R>
__global__ void __launch_bounds__(kCtaThreads, 1) count_kernel(const SweepArgs a0, const typename FE::Params fp, const SweepArgs *batch) {
    const SweepArgs a = batch ? batch[blockIdx.y] : a0;  // (a batch: one sweep per blockIdx.y, general maps)
    using C = Sweep3Cfg<FE, R>;
    constexpr int SEG = C::SEG;
    constexpr uint32_t ALL = C::ALL;
    extern __shared__ __align__(16) unsigned char smem[];
    FE fe;
    fe.init(fp, smem);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (kCtaThreads / 32) + (threadIdx.x >> 5);
    TileWalk<C::TILE_ELEMS> tw;
    tw.init(a, warp, (long long)gridDim.x * (kCtaThreads / 32));
    // the range's carry function, accumulated over all its tiles
    bool t_id = true;
    uint32_t t_const = 0, delta = 0;
    unsigned long long cnt0 = 0;
    for (; tw.tile < tw.end; tw.next(a)) {
        const unsigned long long tile_base = (unsigned long long)tw.tile * C::TILE_ELEMS;
        const bool simple = (tile_base + C::TILE_ELEMS < a.n) && (a.chunk == 0 || tw.ti.rem0 + C::TILE_ELEMS <= a.chunk);
        const bool end_wall = (a.chunk != 0) && (tw.ti.rem0 + C::TILE_ELEMS == a.chunk);
        uint4 w[R];
        uint32_t nx[R];
        load_tile3<FE, R>(a, tile_base, lane, w, nx);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t off = uint32_t(r * C::ROUND_ELEMS + lane * SEG);
            uint32_t next = __shfl_down_sync(FULL, FE::first_elem(w[r]), 1);
            if (lane == 31) next = nx[r];
            uint32_t hv[C::HV], ov[C::HV], valid;
            Walls<SEG> wl;
            const uint32_t m = segment_full<FE, C::TILE_ELEMS>(fe, a, tw.ti, simple, end_wall && r == R - 1 && lane == 31, off,
                                                               tile_base + off, w[r], next, hv, ov, &valid, &wl);
            const uint32_t lead = __clz(~(m << (32 - SEG)));   // ones at the top of the segment
            const uint32_t nid = ~__ballot_sync(FULL, m == ALL);
            const uint32_t cob = __ballot_sync(FULL, (lead & 1u) != 0);
            // carry entering this lane if the range's carry_in is 0
            const uint32_t c_round0 = t_id ? 0u : t_const;
            const uint32_t l_nid = nid & ((1u << lane) - 1);
            const uint32_t cin0 = l_nid ? ((cob >> (31 - __clz(l_nid))) & 1u) : c_round0;
            const uint32_t st = start_bits(m, cin0);
            const uint32_t cnt = __popc(valid & ~((st << 1) | cin0));
            cnt0 += __reduce_add_sync(FULL, cnt);
            if (FE::kMembershipInValue && a.meta != nullptr) {
                // what the walk variant of emit needs to know about this segment before it looks anything up:
                // bit 0 carry_in (if the range's carry_in were 0), bit 1 "carry_in IS the range's carry_in"
                // (everything before it in the range is identity), bit 2 one token fewer if that carry_in is 1,
                // bits 3-6 tokens emitted for carry_in as in bit 0, minus 8 (a full segment emits 8..16),
                // bit 7 every pair of the segment is a rule (identity: it merges all the way at either parity)
                uint32_t dep = 0, d1 = 0;
                if (t_id) {  // warp-uniform, and false for good after the range's first non-identity segment
                    dep = (l_nid == 0u) ? 1u : 0u;
                    const uint32_t st1 = start_bits(m, 1u);
                    d1 = (cnt - __popc(valid & ~((st1 << 1) | 1u))) & 1u;
                }
                const uint32_t c8 = cnt >= 8u ? cnt - 8u : 0u;
                const uint32_t mbyte = cin0 | (dep << 1) | (d1 << 2) | (c8 << 3) | (m == ALL ? 0x80u : 0u);
                if (simple) a.meta[tile_base / SEG + uint32_t(r * 32 + lane)] = static_cast<uint8_t>(mbyte);  // only read for such tiles
            }
            if (nid) {
                if (t_id) {  // the first non-identity segment of the range is the only one whose count sees the range's carry_in
                    const int f = __ffs(nid) - 1;
                    const uint32_t st1 = start_bits(m, 1u);
                    const uint32_t d = cnt - __popc(valid & ~((st1 << 1) | 1u));
                    delta = __shfl_sync(FULL, d, f);
                }
                t_id = false;
                t_const = (cob >> (31 - __clz(nid))) & 1u;
            }
        }
    }
    if (lane == 0 && warp < kMaxRanges) {
        a.scratch.tile_desc[warp] = (t_id ? D_ID : 0u) | (t_const ? D_CONST : 0u) | (delta ? D_DELTA : 0u);
        a.scratch.tile_status[kMaxRanges + warp] = cnt0;
    }
}

// ---------------------------------------------------------------------------------------------------
struct ScanFn {
    uint32_t id, cst, delta;
    unsigned long long cnt0;
};
__device__ __forceinline__ ScanFn scan_compose(const ScanFn &far, const ScanFn &near) {  // carry flows far -> near
    ScanFn r;
    const uint32_t c_mid0 = far.id ? 0u : far.cst;
    r.cnt0 = far.cnt0 + near.cnt0 - ((c_mid0 && near.delta) ? 1u : 0u);
    r.delta = far.id ? near.delta : far.delta;
    r.cst = near.id ? far.cst : near.cst;
    r.id = far.id & near.id;
    return r;
}
__device__ __forceinline__ ScanFn scan_shfl_up(const ScanFn &f, int d) {
    ScanFn o;
    const uint32_t packed = f.id | (f.cst << 1) | (f.delta << 2);
    const uint32_t p = __shfl_up_sync(FULL, packed, d);
    o.id = p & 1u; o.cst = (p >> 1) & 1u; o.delta = (p >> 2) & 1u;
    o.cnt0 = __shfl_up_sync(FULL, f.cnt0, d);
    return o;
}

constexpr int kScanItems = kMaxRanges / kCtaThreads;  // ranges per thread

// One CTA.  tile_status[w] = (carry entering warp w's range) << 63 | tokens emitted by ranges 0..w-1.
__global__ void __launch_bounds__(kCtaThreads, 1) scan_kernel(const SweepArgs a0, int n_ranges, const SweepArgs *batch) {
    const SweepArgs a = batch ? batch[blockIdx.y] : a0;
    __shared__ ScanFn warp_agg[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int t0 = threadIdx.x * kScanItems;
    ScanFn item[kScanItems];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        item[i].id = 1; item[i].cst = 0; item[i].delta = 0; item[i].cnt0 = 0;  // identity padding
        if (t0 + i < n_ranges) {
            const uint32_t d = a.scratch.tile_desc[t0 + i];
            item[i].id = d >> 31; item[i].cst = (d >> 30) & 1u; item[i].delta = (d >> 29) & 1u;
            item[i].cnt0 = a.scratch.tile_status[kMaxRanges + t0 + i];
        }
    }
    ScanFn agg = item[0];
#pragma unroll
    for (int i = 1; i < kScanItems; ++i) agg = scan_compose(agg, item[i]);
    ScanFn inc = agg;  // inclusive scan of the thread aggregates inside the warp, then across the 32 warps
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const ScanFn o = scan_shfl_up(inc, s);
        if (lane >= s) inc = scan_compose(o, inc);
    }
    if (lane == 31) warp_agg[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        ScanFn wa = warp_agg[lane];
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const ScanFn o = scan_shfl_up(wa, s);
            if (lane >= s) wa = scan_compose(o, wa);
        }
        warp_agg[lane] = wa;  // inclusive over warps
    }
    __syncthreads();
    ScanFn ex;  // exclusive prefix function of this thread = (warps before) o (lanes before)
    ex.id = 1; ex.cst = 0; ex.delta = 0; ex.cnt0 = 0;
    if (wid > 0) ex = warp_agg[wid - 1];
    {
        const ScanFn up = scan_shfl_up(inc, 1);
        if (lane > 0) ex = scan_compose(ex, up);
    }
    // the launch starts with carry 0 and nothing emitted
    uint32_t c = ex.id ? 0u : ex.cst;
    unsigned long long b = ex.cnt0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (t0 + i < n_ranges) {
            a.scratch.tile_status[t0 + i] = (c ? R_CARRY : 0ull) | b;
            b += item[i].cnt0 - ((c && item[i].delta) ? 1u : 0u);
            c = item[i].id ? c : item[i].cst;
        }
    }
    if (threadIdx.x == kCtaThreads - 1) {
        *a.scratch.total_tokens = b + a.total_bias;
        *a.scratch.merged_any = (b < a.n || a.total_bias != 0) ? 1u : 0u;
        if (a.out_base_tokens + b > a.out_cap_tokens) *a.scratch.overflow = 1u;
    }
}

// ---------------------------------------------------------------------------------------------------
// One position of the walk variant: if the scan stands here (standing != 0) the table entry of the pair is read
// (it is the big-endian token to emit either way: the merged id, or the element itself when the pair is not a
// rule), stored at sp, and the next position stands unless the pair was a rule.  No branches.
__device__ __forceinline__ void walk_step(uint32_t &sp, uint32_t &standing, uint32_t entry_addr) {
    asm volatile(
        "{\n\t.reg .pred p, h;\n\t.reg .b16 t;\n\t.reg .b32 e;\n\t"
        "setp.ne.u32 p, %1, 0;\n\t"
        "mov.b16 t, 0;\n\t"
        "@p ld.shared.u16 t, [%2];\n\t"
        "@p st.shared.u16 [%0], t;\n\t"
        "@p add.u32 %0, %0, 2;\n\t"
        "cvt.u32.u16 e, t;\n\t"
        "and.b32 e, e, 255;\n\t"
        "setp.ne.u32 h, e, 0;\n\t"
        "selp.u32 %1, 0, 1, h;\n\t}"
        : "+r"(sp), "+r"(standing)
        : "r"(entry_addr)
        : "memory");
}

template <class FE, int R, bool WALK = false>
__global__ void __launch_bounds__(kCtaThreads, 1) emit_kernel(const SweepArgs a0, const typename FE::Params fp, const SweepArgs *batch) {
    const SweepArgs a = batch ? batch[blockIdx.y] : a0;
    using C = Sweep3Cfg<FE, R>;
    constexpr int SEG = C::SEG;
    constexpr int HV = C::HV;
    constexpr uint32_t ALL = C::ALL;
    extern __shared__ __align__(16) unsigned char smem[];
    // the verifying sweep of a general map (tokenizer.rs:83-85): the scan found that nothing merges
    if (a.skip_unmerged_emit && *reinterpret_cast<const volatile uint32_t *>(a.scratch.merged_any) == 0u) return;
    FE fe;
    fe.init(fp, smem);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint16_t *stage = reinterpret_cast<uint16_t *>(smem + FE::TABLE_BYTES) + size_t(threadIdx.x >> 5) * C::STAGE_TOKENS;
    const long long warp = (long long)blockIdx.x * (kCtaThreads / 32) + (threadIdx.x >> 5);
    TileWalk<C::TILE_ELEMS> tw;
    tw.init(a, warp, (long long)gridDim.x * (kCtaThreads / 32));
    if (tw.tile >= tw.end || warp >= kMaxRanges) return;
    const uint64_t rv = a.scratch.tile_status[warp];
    uint32_t carry = uint32_t(rv >> 63);
    const uint32_t range_carry = carry;                      // the carry entering this warp's range
    unsigned long long rel = rv & ~R_CARRY;                  // tokens emitted before the next round (this launch)
    // Output streaming state.  stage[0 .. pend) holds tokens not yet written; stage[0] corresponds to
    // a.out[wpos] and wpos is a multiple of 8 tokens (16 bytes).  The first `head` slots of the very first
    // vector belong to the previous warp's range and must not be written.
    unsigned long long wpos = (rel + a.out_base_tokens) & ~7ull;
    uint32_t pend = uint32_t((rel + a.out_base_tokens) & 7ull);
    uint32_t head = pend;
    // `total` new tokens were appended behind the pending ones: write out the whole 16-byte vectors and keep
    // the < 8 leftover tokens at the front of the staging line for the next round
    auto flush = [&](uint32_t total) {
        const uint32_t have = pend + total;
        const uint32_t nv = have >> 3;
        const bool fits = (wpos + have <= a.out_cap_tokens);
        if (!fits && lane == 0) *a.scratch.overflow = 1u;
        if (fits) {
            for (uint32_t v = lane; v < nv; v += 32) {
                if (v == 0 && head != 0) {
                    for (uint32_t k = head; k < 8; ++k) a.out[wpos + k] = stage[k];
                } else {
                    stg_stream_v4(a.out + wpos + 8 * v, *reinterpret_cast<const uint4 *>(stage + 8 * v));
                }
            }
        }
        const uint32_t rem = have & 7u;
        uint16_t keep = 0;
        if (nv != 0 && lane < rem) keep = stage[8 * nv + lane];
        __syncwarp();
        if (nv != 0) {
            if (lane < rem) stage[lane] = keep;
            head = 0;
            wpos += 8ull * nv;
            pend = rem;
        } else {
            pend = have;
        }
        __syncwarp();
    };
    for (; tw.tile < tw.end; tw.next(a)) {
        const unsigned long long tile_base = (unsigned long long)tw.tile * C::TILE_ELEMS;
        const bool simple = (tile_base + C::TILE_ELEMS < a.n) && (a.chunk == 0 || tw.ti.rem0 + C::TILE_ELEMS <= a.chunk);
        const bool end_wall = (a.chunk != 0) && (tw.ti.rem0 + C::TILE_ELEMS == a.chunk);
        uint4 w[R];
        uint32_t nx[R];
        load_tile3<FE, R>(a, tile_base, lane, w, nx);
        uint32_t mbq[R];
        if constexpr (WALK) {
            if (simple) {
#pragma unroll
                for (int r = 0; r < R; ++r) mbq[r] = a.meta[tile_base / SEG + uint32_t(r * 32 + lane)];
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t off = uint32_t(r * C::ROUND_ELEMS + lane * SEG);
            const unsigned long long g = tile_base + off;
            uint32_t next = __shfl_down_sync(FULL, FE::first_elem(w[r]), 1);
            if (lane == 31) next = nx[r];
            const bool wall_here = end_wall && r == R - 1 && lane == 31;
            if constexpr (WALK) {
            if (simple) {
                // ---- walk variant: the count kernel left carry_in and token count of every segment ----
                const uint32_t mb = mbq[r];
                const uint32_t cin = (mb & 2u) ? range_carry : (mb & 1u);
                const uint32_t cnt = ((mb >> 3) & 15u) + 8u - (cin & (mb >> 2) & (mb >> 1) & 1u);
                uint32_t total = uint32_t(C::ROUND_ELEMS / 2);
                if (__all_sync(FULL, (mb & 0x80u) != 0u)) {
                    // dense warp-round (every lane merges all the way at the carry's parity): 8 lookups, vector store
                    uint32_t hv[HV];
                    fe.lookup_vals(w[r], next, carry, hv);
                    if (pend == 0) {
                        if (wpos + C::ROUND_ELEMS / 2 <= a.out_cap_tokens)
                            stg_stream_v4(a.out + wpos + size_t(lane) * (SEG / 2), make_uint4(hv[0], hv[1], hv[HV > 2 ? 2 : 0], hv[HV > 3 ? 3 : 0]));
                        else if (lane == 0) *a.scratch.overflow = 1u;
                        wpos += C::ROUND_ELEMS / 2;
                    } else {
                        uint16_t *d = stage + pend + lane * (SEG / 2);
                        if ((pend & 1u) == 0) {
#pragma unroll
                            for (int k = 0; k < HV; ++k) reinterpret_cast<uint32_t *>(d)[k] = hv[k];
                        } else {
#pragma unroll
                            for (int k = 0; k < HV; ++k) { d[2 * k] = uint16_t(hv[k]); d[2 * k + 1] = uint16_t(hv[k] >> 16); }
                        }
                        __syncwarp();
                        flush(total);
                    }
                } else {
                    uint32_t incl = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(FULL, incl, d);
                        if (lane >= d) incl += t;
                    }
                    const uint32_t pos = incl - cnt;
                    total = __shfl_sync(FULL, incl, 31);
                    // the scan walks the lane's 16 positions; only where it stands is the table read
                    uint32_t sp = uint32_t(__cvta_generic_to_shared(stage + pend + pos));
                    const uint32_t tb = uint32_t(__cvta_generic_to_shared(fe.tbl));
                    uint32_t standing = cin ^ 1u;
                    const uint32_t wd[5] = {w[r].x, w[r].y, w[r].z, w[r].w, next};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t we = wd[k], wo = __funnelshift_r(wd[k], wd[k + 1], 8);   // pairs at 4k,4k+2 / 4k+1,4k+3
                        const uint32_t ye = we ^ ((we >> 7) & 0x01FF01FFu), yo = wo ^ ((wo >> 7) & 0x01FF01FFu);
                        walk_step(sp, standing, tb + ((ye << 1) & 0x1FFFEu));
                        walk_step(sp, standing, tb + ((yo << 1) & 0x1FFFEu));
                        walk_step(sp, standing, tb + ((ye >> 15) & 0x1FFFEu));
                        if (k == 3 && wall_here) {  // the chunk's last element: never a pair, emitted as it is
                            if (standing) {
                                *reinterpret_cast<uint16_t *>(__cvta_shared_to_generic(sp)) = uint16_t(FE::raw_be(w[r], SEG - 1));
                                sp += 2;
                            }
                            standing = 1u;
                        } else {
                            walk_step(sp, standing, tb + ((yo >> 15) & 0x1FFFEu));
                        }
                    }
                    __syncwarp();
                    flush(total);
                    carry = __shfl_sync(FULL, standing ^ 1u, 31);
                }
                rel += total;
                if (wall_here && a.chunk_ends != nullptr) a.chunk_ends[tw.ti.ck0] = a.chunk_ends_base + 2ull * rel;
                continue;
            }
            }
            // ---- dense warp-round: the pairs of the carry's parity are all rules ----
            uint32_t hv[HV], ov[HV], valid;
            int have_par = -1;
            if (FE::kMembershipInValue && simple) {
                fe.lookup_vals(w[r], next, carry, hv);
                have_par = int(carry);
                const bool ok = FE::all_present(hv) && !(wall_here && carry);
                if (__all_sync(FULL, ok)) {
                    if (pend == 0) {  // output is vector-aligned: straight from registers
                        if (wpos + C::ROUND_ELEMS / 2 <= a.out_cap_tokens) {
                            uint16_t *dst = a.out + wpos + size_t(lane) * (SEG / 2);
                            if (SEG == 16) stg_stream_v4(dst, make_uint4(hv[0], hv[1], hv[HV > 2 ? 2 : 0], hv[HV > 3 ? 3 : 0]));
                            else *reinterpret_cast<uint2 *>(dst) = make_uint2(hv[0], hv[1]);
                        } else if (lane == 0) {
                            *a.scratch.overflow = 1u;
                        }
                        wpos += C::ROUND_ELEMS / 2;
                    } else {          // an earlier break shifted the output: go through the staging line
                        uint16_t *d = stage + pend + lane * (SEG / 2);
                        if ((pend & 1u) == 0) {
#pragma unroll
                            for (int k = 0; k < HV; ++k) reinterpret_cast<uint32_t *>(d)[k] = hv[k];
                        } else {
#pragma unroll
                            for (int k = 0; k < HV; ++k) { d[2 * k] = uint16_t(hv[k]); d[2 * k + 1] = uint16_t(hv[k] >> 16); }
                        }
                        __syncwarp();
                        flush(uint32_t(C::ROUND_ELEMS / 2));
                    }
                    rel += C::ROUND_ELEMS / 2;
                    if (wall_here && a.chunk_ends != nullptr) a.chunk_ends[tw.ti.ck0] = a.chunk_ends_base + 2ull * rel;
                    continue;  // carry is unchanged
                }
            }
            // ---- general warp-round: the other parity too (the first one is reused), full membership word ----
            Walls<SEG> wl;
            const uint32_t m = segment_full<FE, C::TILE_ELEMS>(fe, a, tw.ti, simple, wall_here, off, g, w[r], next, hv, ov, &valid, &wl,
                                                               have_par);
            const uint32_t lead = __clz(~(m << (32 - SEG)));
            const uint32_t nid = ~__ballot_sync(FULL, m == ALL);
            const uint32_t cob = __ballot_sync(FULL, (lead & 1u) != 0);
            const uint32_t l_nid = nid & ((1u << lane) - 1);
            const uint32_t cin = l_nid ? ((cob >> (31 - __clz(l_nid))) & 1u) : carry;
            const uint32_t st = start_bits(m, cin);
            const uint32_t em = valid & ~((st << 1) | cin);
            const uint32_t cnt = __popc(em);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += t;
            }
            const uint32_t pos = incl - cnt;
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            if (a.chunk_ends != nullptr && wl.endm) {
                uint32_t e = wl.endm;
                unsigned long long ck = wl.ck;
                while (e) {
                    const int d = __ffs(e) - 1;
                    e &= e - 1;
                    a.chunk_ends[ck++] = a.chunk_ends_base + 2ull * (rel + pos + __popc(em & ((2u << d) - 1)));
                }
            }
            // one predicated 2-byte shared store + one predicated pointer bump per position
            uint32_t sp = uint32_t(__cvta_generic_to_shared(stage + pend + pos));
#pragma unroll
            for (int j = 0; j < SEG; ++j) {
                const uint32_t v = (j & 1) ? ov[j >> 2] : hv[j >> 2];
                const uint32_t tok = ((j >> 1) & 1) ? (v >> 16) : v;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t.reg .b16 t;\n\t"
                    "setp.ne.u32 p, %2, 0;\n\t"
                    "cvt.u16.u32 t, %1;\n\t"
                    "@p st.shared.u16 [%0], t;\n\t"
                    "@p add.u32 %0, %0, 2;\n\t}"
                    : "+r"(sp)
                    : "r"(tok), "r"(em & (1u << j))
                    : "memory");
            }
            __syncwarp();
            flush(total);
            rel += total;
            if (nid) carry = (cob >> (31 - __clz(nid))) & 1u;
        }
    }
    // tail of the range: the leftover tokens (the next warp's range starts right behind them)
    if (pend > head && lane >= head && lane < pend) {
        if (wpos + pend <= a.out_cap_tokens) a.out[wpos + lane] = stage[lane];
        else *a.scratch.overflow = 1u;
    }
}

template <class FE, int R, bool WALK = false>
struct Sweep3Launch {
    using C = Sweep3Cfg<FE, R>;
    static constexpr size_t smem_count = FE::TABLE_BYTES;
    static constexpr size_t smem_emit = FE::TABLE_BYTES + size_t(kCtaThreads / 32) * C::STAGE_TOKENS * 2;
    static_assert(smem_emit <= 227 * 1024, "shared memory budget");
    // opt the two big kernels into their dynamic shared memory (once per device; also what a device-side
    // launch of them relies on)
    static cudaError_t configure(int dev) {
        static std::atomic<bool> configured[kMaxDevices];
        if (dev >= kMaxDevices) return cudaErrorInvalidDevice;
        if (!configured[dev].load(std::memory_order_acquire)) {
            cudaError_t err = cudaFuncSetAttribute(count_kernel<FE, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_count));
            if (err != cudaSuccess) return err;
            err = cudaFuncSetAttribute(emit_kernel<FE, R, WALK>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_emit));
            if (err != cudaSuccess) return err;
            configured[dev].store(true, std::memory_order_release);
        }
        return cudaSuccess;
    }
    static unsigned grid_for(size_t n, int dev) {
        const size_t n_tiles = (n + C::TILE_ELEMS - 1) / C::TILE_ELEMS;
        const size_t warps_per_cta = kCtaThreads / 32;
        size_t grid = (n_tiles + warps_per_cta - 1) / warps_per_cta;
        if (grid > size_t(sm_count(dev))) grid = size_t(sm_count(dev));
        if (grid > size_t(kMaxRanges) / warps_per_cta) grid = size_t(kMaxRanges) / warps_per_cta;
        if (grid == 0) grid = 1;
        return unsigned(grid);
    }
};

// Host launch: control block cleared, then count, scan, emit on `stream`.
template <class FE, int R, bool WALK = false>
cudaError_t launch_sweep3(const SweepArgs &a_in, const typename FE::Params &fp, cudaStream_t stream) {
    using L = Sweep3Launch<FE, R, WALK>;
    SweepArgs a = a_in;
    a.meta = WALK ? a.scratch.meta : nullptr;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    err = L::configure(dev);
    if (err != cudaSuccess) return err;
    if (a.scratch.max_tiles < size_t(2 * kMaxRanges)) return cudaErrorInvalidValue;
    err = cudaMemsetAsync(a.scratch.ctrl, 0, 256, stream);
    if (err != cudaSuccess) return err;
    const unsigned grid = L::grid_for(a.n, dev);
    count_kernel<FE, R><<<dim3(grid), dim3(kCtaThreads), L::smem_count, stream>>>(a, fp, nullptr);
    scan_kernel<<<1, kCtaThreads, 0, stream>>>(a, int(grid * (kCtaThreads / 32)), nullptr);
    emit_kernel<FE, R, WALK><<<dim3(grid), dim3(kCtaThreads), L::smem_emit, stream>>>(a, fp, nullptr);
    return cudaGetLastError();
}

// A batch of independent sweeps (the chunks of a general map's sweep level) in three launches: blockIdx.y picks the
// sweep, whose arguments are read from `d_args` (device-readable, e.g. page-locked host memory; `h_args` is the host's
// view of the same array).  The front-end tables are loaded once per CTA for 1/grid.x of a sweep instead of once per
// CTA and chunk, and a level costs three launches instead of three per chunk.
template <class FE, int R>
cudaError_t launch_sweep3_batch(const SweepArgs *h_args, const SweepArgs *d_args, int nb, const typename FE::Params &fp, cudaStream_t stream) {
    using L = Sweep3Launch<FE, R, false>;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    err = L::configure(dev);
    if (err != cudaSuccess) return err;
    size_t n_max = 0;
    for (int j = 0; j < nb; ++j) {
        if (h_args[j].scratch.max_tiles < size_t(2 * kMaxRanges) || h_args[j].meta != nullptr) return cudaErrorInvalidValue;
        err = cudaMemsetAsync(h_args[j].scratch.ctrl, 0, 256, stream);
        if (err != cudaSuccess) return err;
        n_max = std::max(n_max, h_args[j].n);
    }
    unsigned grid = L::grid_for(n_max, dev);
    const unsigned share = std::max(1u, unsigned(sm_count(dev)) / unsigned(nb));  // about one CTA per SM over the whole batch
    if (grid > share) grid = share;
    count_kernel<FE, R><<<dim3(grid, nb), dim3(kCtaThreads), L::smem_count, stream>>>(h_args[0], fp, d_args);
    scan_kernel<<<dim3(1, nb), kCtaThreads, 0, stream>>>(h_args[0], int(grid * (kCtaThreads / 32)), d_args);
    emit_kernel<FE, R, false><<<dim3(grid, nb), dim3(kCtaThreads), L::smem_emit, stream>>>(h_args[0], fp, d_args);
    return cudaGetLastError();
}

// Device launch (CUDA dynamic parallelism, tail-launch stream): the same three kernels, enqueued by the dense
// pass when its speculation failed.  They start when the launching grid has completed and run in order.
// The caller has cleared the control block.  Returns false if a launch was refused.
template <class FE, int R, bool WALK = false>
__device__ bool tail_launch_sweep3(const SweepArgs &a_in, const typename FE::Params &fp, unsigned grid) {
    using L = Sweep3Launch<FE, R, WALK>;
    SweepArgs a = a_in;
    a.meta = WALK ? a.scratch.meta : nullptr;
    count_kernel<FE, R><<<dim3(grid), dim3(kCtaThreads), L::smem_count, cudaStreamTailLaunch>>>(a, fp, nullptr);
    scan_kernel<<<1, kCtaThreads, 0, cudaStreamTailLaunch>>>(a, int(grid * (kCtaThreads / 32)), nullptr);
    emit_kernel<FE, R, WALK><<<dim3(grid), dim3(kCtaThreads), L::smem_emit, cudaStreamTailLaunch>>>(a, fp, nullptr);
    return cudaGetLastError() == cudaSuccess;
}
