// fused.cuh -- the exact sweep in ONE pass: count, decoupled look-back and emit fused in a single persistent kernel.
//
// The input is read from DRAM once (N_in + 2*T_out bytes of traffic = the algorithmic bytes) and every pair is
// looked up once.  One CTA per SM holds the 128 KiB pair table; its warps form NG independent groups of WG warps.
// A group takes tiles of WG*R*512 input bytes in ticket order (one global counter):
//
//   stage   the tile is copied into the group's shared-memory buffer by cp.async.bulk (one elected lane, one
//           mbarrier per group), issued while the group still emits the previous tile;
//   count   every warp owns R*512 consecutive bytes of the tile (R rounds of 32 lanes x 16 bytes): both parities
//           are looked up, the tokens stay in registers, run parity is resolved inside the warp with two ballots
//           per round under the hypothesis "the warp's carry_in is 0" and the warp's slice is reduced to one carry
//           function (identity / constant, tokens for carry_in 0, the 0/1-token delta for carry_in 1);
//   chain   the group's first warp composes the WG warp functions, publishes the tile's function (status A) in
//           the tile's 64-bit descriptor and polls the 64 descriptors in front of it until it sees an inclusive
//           prefix (status P) with nothing missing behind it.  The carry entering every tile of the window comes
//           from two ballots (nearest non-identity tile in front of it), its exact token count is then
//           cnt0 - (delta & carry_in), and one warp-wide add gives the offset: the chain advances up to 64 tiles
//           per hop (poll latency + ~40 instructions + store visibility).  Chunk walls lie on tile boundaries, where
//           the carry is 0 by definition; chunk_ends fall out of the inclusive prefixes;
//   emit    every warp compacts its retained tokens into a warp-private staging line (XOR-swizzled so that the
//           32 lanes' 2-byte stores spread over the banks) and streams whole 4-byte words out.  Only the lanes in
//           front of the slice's first non-identity segment depend on the carry_in; they are redone when it is 1.
//
// While one group waits for its look-back the other groups of the CTA keep the SM busy.  Forward progress: a tile
// only ever waits for tiles with smaller tickets, which are held by resident groups.
// Included by kernels.cu inside its anonymous namespace, after sweep3.cuh (ScanFn, scan_compose, start_bits).
#pragma once

constexpr unsigned long long FZ_A = 1ull << 62, FZ_P = 2ull << 62;
constexpr unsigned long long FZ_A_ID = 1ull << 61, FZ_A_CST = 1ull << 60, FZ_A_DELTA = 1ull << 59;
constexpr unsigned long long FZ_P_CARRY = 1ull << 60;  // same bit as FZ_A_CST: "the carry leaving this tile"
constexpr unsigned long long FZ_COUNT = (1ull << 56) - 1;
constexpr uint32_t FZ_F_WALL = 1u, FZ_F_START = 2u;

struct FusedGroupShared {
    unsigned long long mbar;        // completion of the bulk copy into the group's buffer
    uint32_t next_tile;             // ticket of the group's next tile (written by the leader between the two barriers)
    uint32_t next_flags;            // FZ_F_WALL: its last element is chunk-last; FZ_F_START: it starts a chunk
    unsigned long long fn_cnt[16];  // per warp: tokens of its slice for carry_in 0
    uint32_t fn_flags[16];          // per warp: bit0 identity, bit1 constant carry_out, bit2 delta
    unsigned long long res[16];     // per warp: carry_in << 63 | tokens of the launch in front of its slice
};

template <int NG, int WG, int R>
struct FusedCfg {
    static_assert(WG <= 16, "the leader scans the warp functions in one half warp");
    static constexpr int THREADS = NG * WG * 32;
    static constexpr int WARP_BYTES = R * 512;
    static constexpr int TILE = WG * WARP_BYTES;
    static constexpr int BUF = TILE + 128;     // + the look-ahead vector; keeps every buffer 128-byte aligned
    static constexpr int STAGE_BYTES = 1152;   // per warp: 1 pending + 512 new tokens, in whole 128-byte swizzle windows
    static constexpr int OFF_STAGE = PairsFE::TABLE_BYTES;
    static constexpr int OFF_BUF = OFF_STAGE + NG * WG * STAGE_BYTES;
    static constexpr int OFF_GS = OFF_BUF + NG * BUF;
    static constexpr int GS_BYTES = 512;
    static_assert(sizeof(FusedGroupShared) <= GS_BYTES, "group block");
    static constexpr int SMEM = OFF_GS + NG * GS_BYTES;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// false: the copy did not land within ~2 s (reported as a CUDA error by the host instead of hanging the device)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (int tries = 0; tries < 4096 && ok == 0u; ++tries) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x80000;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
    return ok != 0u;
}
// global -> shared bulk copy (the TMA unit's 1-D form): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void group_bar(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void stg_stream_u32(void *p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// staging line: shared address of a token or word -> where it really lives.  The bank bits are XORed with the index
// of the 128-byte window, so that stores 4 to 8 words apart (one lane's tokens behind the other's) do not pile up on
// a few banks.  Every window is permuted within itself: lines are whole, 128-byte aligned windows.
__device__ __forceinline__ uint32_t stage_swz(uint32_t addr) { return addr ^ ((addr >> 5) & 0x7Cu); }

// membership word of one 16-byte segment from the looked-up tokens (present <=> low byte != 0): bit j <-> position j
__device__ __forceinline__ uint32_t fz_membership(const uint32_t *hv, const uint32_t *ov) {
    uint32_t p[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t x = __byte_perm(hv[k], ov[k], 0x6240);                       // low bytes of positions 4k .. 4k+3
        const uint32_t f = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;  // bit 7 of a byte: it is non-zero
        p[k] = f * 0x00204081u;                                                     // the four flags land in bits 28 .. 31
    }
    return (p[0] >> 28) | ((p[1] >> 24) & 0xF0u) | ((p[2] >> 20) & 0xF00u) | ((p[3] >> 16) & 0xF000u);
}

// Leader lane: starts the copy of tile `t` into the group's buffer.  Whole 16-byte vectors go through the bulk copy
// (tile + look-ahead vector where the input has them); the < 16 ragged bytes of the input's end are left to
// fz_copy_tail, which the whole leader warp runs.
template <class C>
__device__ __forceinline__ void fz_issue_copy(const SweepArgs &a, uint32_t t, unsigned char *buf, uint32_t bar) {
    const unsigned long long base = (unsigned long long)t * C::TILE;
    const unsigned long long left = a.n - base;
    const uint32_t avail = left < (unsigned long long)(C::TILE + 16) ? uint32_t(left) : uint32_t(C::TILE + 16);
    const uint32_t bytes16 = avail & ~15u;
    if (bytes16 != 0) {
        mbar_expect_tx(bar, bytes16);
        bulk_g2s(smem_u32(buf), static_cast<const unsigned char *>(a.in) + base, bytes16, bar);
    } else {
        mbar_arrive(bar);
    }
}
template <class C>
__device__ __forceinline__ void fz_copy_tail(const SweepArgs &a, uint32_t t, unsigned char *buf, int lane) {
    const unsigned long long base = (unsigned long long)t * C::TILE;
    const unsigned long long left = a.n - base;
    if (left >= (unsigned long long)(C::TILE + 16)) return;
    const uint32_t avail = uint32_t(left);
    const uint32_t bytes16 = avail & ~15u;
    if (bytes16 + lane < avail) buf[bytes16 + lane] = static_cast<const unsigned char *>(a.in)[base + bytes16 + lane];
}
// walls lie on tile boundaries: tpc = tiles per chunk (0: the whole input is one chunk)
__device__ __forceinline__ uint32_t fz_tile_flags(uint32_t t, uint32_t n_tiles, uint32_t tpc) {
    uint32_t f = 0;
    if (t + 1 == n_tiles || (tpc != 0 && (t + 1) % tpc == 0)) f |= FZ_F_WALL;
    if (t == 0 || (tpc != 0 && t % tpc == 0)) f |= FZ_F_START;
    return f;
}

template <int NG, int WG, int R>
__global__ void __launch_bounds__(NG *WG * 32, 1)
fused_sweep_kernel(const SweepArgs a, const uint16_t *__restrict__ table, unsigned long long *__restrict__ desc,
                   uint32_t *__restrict__ tile_counter, uint32_t n_tiles, uint32_t tpc) {
    using C = FusedCfg<NG, WG, R>;
    constexpr int GT = WG * 32;  // threads of a group
    extern __shared__ __align__(16) unsigned char smem[];
    PairsFE fe;
    PairsFE::Params fp{table};
    fe.init(fp, smem);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int g = warp / WG, wg = warp % WG;
    const bool leader = (wg == 0);
    FusedGroupShared *gs = reinterpret_cast<FusedGroupShared *>(smem + C::OFF_GS + g * C::GS_BYTES);
    unsigned char *buf = smem + C::OFF_BUF + g * C::BUF;
    const uint32_t stage_s = smem_u32(smem + C::OFF_STAGE + warp * C::STAGE_BYTES);
    const uint32_t bar = smem_u32(&gs->mbar);

    // ---- prologue: barrier, first ticket, first copy --------------------------------------------------
    if (leader) {
        if (lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        uint32_t first = 0;
        if (lane == 0) {
            first = atomicAdd(tile_counter, 1u);
            gs->next_tile = first;
            gs->next_flags = fz_tile_flags(first, n_tiles, tpc);
            if (first < n_tiles) fz_issue_copy<C>(a, first, buf, bar);
        }
        first = __shfl_sync(FULL, first, 0);
        if (first < n_tiles) fz_copy_tail<C>(a, first, buf, lane);
    }
    __syncthreads();  // table, barriers, first tickets
    uint32_t cur = gs->next_tile;
    uint32_t flags = gs->next_flags;
    uint32_t parity = 0;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t lane4 = uint32_t(lane) << 2;

    while (cur < n_tiles) {
        uint32_t nxt = 0;
        if (leader && lane == 0) nxt = atomicAdd(tile_counter, 1u);  // consumed after the count phase
        const bool wall_end = (flags & FZ_F_WALL) != 0;
        const bool full = (cur + 1 != n_tiles) || (a.n % C::TILE == 0);
        const uint32_t tile_len = full ? uint32_t(C::TILE) : uint32_t(a.n % C::TILE);
        if (!mbar_wait(bar, parity)) *a.scratch.overflow = 3u;
        parity ^= 1u;

        // ---- count: lookups (retained), run parity under carry_in = 0, the slice's carry function ----
        uint32_t hv[R][4], ov[R][4];
        uint32_t emw[R];  // bits 0-15: positions emitted
        bool t_id = true;
        uint32_t t_const = 0, delta = 0, cnt0 = 0;
        const unsigned char *slice = buf + wg * C::WARP_BYTES;
        const uint32_t slice_off = uint32_t(wg * C::WARP_BYTES);
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const uint32_t off = slice_off + uint32_t(k * 512 + lane * 16);  // offset of the lane's segment in the tile
            const uint4 w = *reinterpret_cast<const uint4 *>(slice + k * 512 + lane * 16);
            uint32_t next = __shfl_down_sync(FULL, w.x & 0xffu, 1);
            if (lane == 31) next = slice[k * 512 + 512];
            fe.lookup_vals(w, next, 0u, hv[k]);
            fe.lookup_vals(w, next, 1u, ov[k]);
            uint32_t valid = 0xFFFFu;
            if (!full) valid = (off + 16 <= tile_len) ? 0xFFFFu : (off < tile_len ? ((1u << (tile_len - off)) - 1u) : 0u);
            if (wall_end && (!full || (k == R - 1 && wg == WG - 1))) {  // warp-uniform
                // the wall suppresses the pair that starts at the tile's last element: the raw token goes out there
                const uint32_t dj = tile_len - 1u - off;  // >= 16 (or wrapped) in every segment but one
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (dj == uint32_t(j)) {
                        const uint32_t be = PairsFE::raw_be(w, j);
                        uint32_t &dst = (j & 1) ? ov[k][j >> 2] : hv[k][j >> 2];
                        dst = ((j >> 1) & 1) ? ((dst & 0x0000ffffu) | (be << 16)) : ((dst & 0xffff0000u) | be);
                    }
                }
            }
            const uint32_t m = fz_membership(hv[k], ov[k]) & valid;
            const uint32_t lead = __clz(~(m << 16));  // ones at the top of the segment
            const uint32_t nid = ~__ballot_sync(FULL, m == 0xFFFFu);
            const uint32_t cob = __ballot_sync(FULL, (lead & 1u) != 0);
            const uint32_t c_round0 = t_id ? 0u : t_const;
            const uint32_t l_nid = nid & lt_mask;
            const uint32_t cin0 = l_nid ? ((cob >> (31 - __clz(l_nid))) & 1u) : c_round0;
            const uint32_t st = start_bits(m, cin0);
            const uint32_t em = valid & ~((st << 1) | cin0);
            const uint32_t cnt = __popc(em);
            emw[k] = em;
            cnt0 += __reduce_add_sync(FULL, cnt);
            if (nid) {
                if (t_id) {  // the slice's first non-identity segment is the only one whose count sees the slice's carry_in
                    const int f = __ffs(nid) - 1;
                    const uint32_t st1 = start_bits(m, 1u);
                    const uint32_t d = cnt - __popc(valid & ~((st1 << 1) | 1u));
                    delta = __shfl_sync(FULL, d, f);
                }
                t_id = false;
                t_const = (cob >> (31 - __clz(nid))) & 1u;
            }
        }
        if (lane == 0) {
            gs->fn_flags[wg] = (t_id ? 1u : 0u) | (t_const << 1) | (delta << 2);
            gs->fn_cnt[wg] = cnt0;
        }
        group_bar(1 + g, GT);

        // ---- chain: the group's first warp scans the warp functions and resolves the tile's prefix ----
        if (leader) {
            ScanFn item;
            item.id = 1; item.cst = 0; item.delta = 0; item.cnt0 = 0;
            if (lane < WG) {
                const uint32_t fl = gs->fn_flags[lane];
                item.id = fl & 1u; item.cst = (fl >> 1) & 1u; item.delta = (fl >> 2) & 1u;
                item.cnt0 = gs->fn_cnt[lane];
            }
            ScanFn inc = item;
#pragma unroll
            for (int s = 1; s < WG; s <<= 1) {
                const ScanFn o = scan_shfl_up(inc, s);
                if (lane >= s) inc = scan_compose(o, inc);
            }
            ScanFn ex = scan_shfl_up(inc, 1);
            if (lane == 0) { ex.id = 1; ex.cst = 0; ex.delta = 0; ex.cnt0 = 0; }
            ScanFn tf;  // the whole tile
            {
                const uint32_t packed = inc.id | (inc.cst << 1) | (inc.delta << 2);
                const uint32_t p = __shfl_sync(FULL, packed, WG - 1);
                tf.id = p & 1u; tf.cst = (p >> 1) & 1u; tf.delta = (p >> 2) & 1u;
                tf.cnt0 = __shfl_sync(FULL, inc.cnt0, WG - 1);
            }
            const bool starts = (flags & FZ_F_START) != 0;
            if (starts) {  // the carry entering a chunk is 0: the tile's function collapses to a constant
                tf.cst = tf.id ? 0u : tf.cst;
                tf.id = 0; tf.delta = 0;
            }
            if (lane == 0) {
                st_desc(desc + cur, FZ_A | (tf.id ? FZ_A_ID : 0ull) | (tf.cst ? FZ_A_CST : 0ull) | (tf.delta ? FZ_A_DELTA : 0ull) | tf.cnt0);
                // the buffer has been read by every warp of the group: the next tile may land in it
                gs->next_tile = nxt;
                gs->next_flags = fz_tile_flags(nxt, n_tiles, tpc);
                if (nxt < n_tiles) fz_issue_copy<C>(a, nxt, buf, bar);
            }
            nxt = __shfl_sync(FULL, nxt, 0);
            if (nxt < n_tiles) fz_copy_tail<C>(a, nxt, buf, lane);
            // look-back over the 64 tiles in front: position x <-> tile cur-1-x; lane i holds positions i and i+32
            uint32_t c_in = 0;
            unsigned long long base = 0;
            {
                const long long i0 = (long long)cur - 1 - lane, i1 = i0 - 32;
                unsigned long long d0 = FZ_P, d1 = FZ_P;  // in front of tile 0: carry 0, nothing emitted
                uint32_t polls = 0;
                for (;;) {
                    if (i0 >= 0) d0 = ld_desc(desc + i0);
                    if (i1 >= 0) d1 = ld_desc(desc + i1);
                    const uint32_t s0 = uint32_t(d0 >> 62), s1 = uint32_t(d1 >> 62);
                    const unsigned long long pm =
                        (unsigned long long)__ballot_sync(FULL, s0 == 2u) | ((unsigned long long)__ballot_sync(FULL, s1 == 2u) << 32);
                    const unsigned long long zm =
                        (unsigned long long)__ballot_sync(FULL, s0 == 0u) | ((unsigned long long)__ballot_sync(FULL, s1 == 0u) << 32);
                    if (pm != 0ull) {
                        const int q = __ffsll((long long)pm) - 1;  // the nearest inclusive prefix
                        const unsigned long long nearer = (1ull << q) - 1ull;
                        if ((zm & nearer) == 0ull) {
                            // the carry leaving position x: P.carry at q, the constant of a non-identity tile, else whatever enters it
                            unsigned long long nim = (unsigned long long)__ballot_sync(FULL, (d0 & FZ_A_ID) == 0ull) |
                                                     ((unsigned long long)__ballot_sync(FULL, (d1 & FZ_A_ID) == 0ull) << 32);
                            const unsigned long long cm = (unsigned long long)__ballot_sync(FULL, (d0 & FZ_A_CST) != 0ull) |
                                                          ((unsigned long long)__ballot_sync(FULL, (d1 & FZ_A_CST) != 0ull) << 32);
                            nim = (nim & nearer) | (1ull << q);
                            // the carry entering position x leaves the nearest non-identity position behind it (x+1 .. q)
                            const unsigned long long above0 = nim >> (lane + 1);
                            const int j0 = lane + __ffsll((long long)above0);
                            const uint32_t cin0 = uint32_t((cm >> (j0 & 63)) & 1ull);
                            const unsigned long long above1 = (lane < 31) ? (nim >> (lane + 33)) : 0ull;
                            const int j1 = lane + 32 + __ffsll((long long)above1);
                            const uint32_t cin1 = uint32_t((cm >> (j1 & 63)) & 1ull);
                            uint32_t e = 0;
                            if (lane < q) e += uint32_t(d0 & 0xffffffffull) - (((d0 & FZ_A_DELTA) && cin0) ? 1u : 0u);
                            if (lane + 32 < q) e += uint32_t(d1 & 0xffffffffull) - (((d1 & FZ_A_DELTA) && cin1) ? 1u : 0u);
                            e = __reduce_add_sync(FULL, e);
                            const unsigned long long pd0 = __shfl_sync(FULL, d0, q & 31), pd1 = __shfl_sync(FULL, d1, q & 31);
                            base = ((q < 32 ? pd0 : pd1) & FZ_COUNT) + e;
                            const unsigned long long low = nim & (0ull - nim);  // the nearest non-identity position (q at the latest)
                            c_in = (cm & low) ? 1u : 0u;                        // its carry enters this tile
                            break;
                        }
                    }
                    if (++polls == (1u << 22)) {  // seconds: a predecessor died; fail the launch instead of hanging
                        *a.scratch.overflow = 3u;
                        break;
                    }
                    __nanosleep(40);
                }
            }
            if (starts) c_in = 0;
            const uint32_t c_out = tf.id ? c_in : tf.cst;
            const unsigned long long total = base + tf.cnt0 - ((c_in && tf.delta) ? 1ull : 0ull);
            if (lane == 0) {
                st_desc(desc + cur, FZ_P | (c_out ? FZ_P_CARRY : 0ull) | total);
                if (wall_end && a.chunk_ends != nullptr) a.chunk_ends[tpc ? cur / tpc : 0u] = a.chunk_ends_base + 2ull * total;
                if (cur == n_tiles - 1) {
                    *a.scratch.total_tokens = total;
                    *a.scratch.merged_any = (total < a.n) ? 1u : 0u;
                    if (a.out_base_tokens + total > a.out_cap_tokens) *a.scratch.overflow = 1u;
                }
            }
            if (lane < WG) {
                // (the scan ran on the warp functions as they are: a chunk start only fixes the carry entering warp 0)
                const uint32_t cw = ex.id ? c_in : ex.cst;
                const unsigned long long bw = base + ex.cnt0 - ((c_in && ex.delta) ? 1ull : 0ull);
                gs->res[lane] = (cw ? R_CARRY : 0ull) | bw;
            }
        }
        group_bar(1 + g, GT);

        // ---- emit: compaction of the retained tokens, streamed out in whole words -----------------------
        {
            const unsigned long long rv = gs->res[wg];
            cur = gs->next_tile;
            flags = gs->next_flags;
            const uint32_t slice_carry = uint32_t(rv >> 63);
            const unsigned long long abs0 = (rv & ~R_CARRY) + a.out_base_tokens;
            // stage[0 .. pend) holds tokens not yet written; logical token 0 of the line corresponds to a.out[wpos],
            // wpos is even.  `head`: that slot belongs to the slice in front of this one and is not written here.
            unsigned long long wpos = abs0 & ~1ull;
            uint32_t pend = uint32_t(abs0 & 1ull);
            uint32_t head = pend;
            bool dep = slice_carry != 0u;  // the lanes in front of the slice's first non-identity segment see carry_in = 1
            auto flush = [&](uint32_t total) {
                const uint32_t have = pend + total;
                const uint32_t nw = have >> 1;
                const bool fits = (wpos + have <= a.out_cap_tokens);
                if (!fits && lane == 0) *a.scratch.overflow = 1u;
                if (fits && nw != 0) {
                    unsigned char *gout = reinterpret_cast<unsigned char *>(a.out + wpos) + lane4;
                    {  // word v = lane + 32 i lives in window i of the line
                        const uint32_t word = lds_u32(stage_swz(stage_s + lane4));
                        if (uint32_t(lane) < nw) {
                            if (lane == 0 && head != 0) *reinterpret_cast<uint16_t *>(gout + 2) = uint16_t(word >> 16);
                            else stg_stream_u32(gout, word);
                        }
                    }
#pragma unroll
                    for (int i = 1; i < 8; ++i) {
                        if (uint32_t(32 * i) < nw) {  // warp-uniform
                            const uint32_t word = lds_u32(stage_swz(stage_s + 128u * i + lane4));
                            if (uint32_t(lane + 32 * i) < nw) stg_stream_u32(gout + 128 * i, word);
                        }
                    }
                }
                uint32_t keep = 0;
                const bool odd = (have & 1u) != 0;
                if (nw != 0 && odd && lane == 0)
                    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(keep) : "r"(stage_swz(stage_s + 2u * (have - 1u))) : "memory");
                __syncwarp();
                if (nw != 0) {
                    if (odd && lane == 0) asm volatile("st.shared.u16 [%0], %1;" ::"r"(stage_swz(stage_s)), "h"(uint16_t(keep)) : "memory");
                    head = 0;
                    wpos += 2ull * nw;
                    pend = have & 1u;
                } else {
                    pend = have;
                }
                __syncwarp();
            };
#pragma unroll
            for (int k = 0; k < R; ++k) {
                uint32_t em = emw[k];
                if (dep) {  // warp-uniform; false for good after the slice's first non-identity segment
                    const uint32_t off = slice_off + uint32_t(k * 512 + lane * 16);
                    uint32_t valid = 0xFFFFu;
                    if (!full) valid = (off + 16 <= tile_len) ? 0xFFFFu : (off < tile_len ? ((1u << (tile_len - off)) - 1u) : 0u);
                    const uint32_t m = fz_membership(hv[k], ov[k]) & valid;
                    const uint32_t nid = ~__ballot_sync(FULL, m == 0xFFFFu);
                    if ((nid & lt_mask) == 0u) {
                        const uint32_t st1 = start_bits(m, 1u);
                        em = valid & ~((st1 << 1) | 1u);
                    }
                    if (nid) dep = false;
                }
                const uint32_t x = em ^ 0x5555u;
                const bool dense0 = __all_sync(FULL, x == 0u), dense1 = __all_sync(FULL, x == 0xFFFFu);
                if ((dense0 || dense1) && pend == 0 && (wpos & 7ull) == 0) {
                    // every lane emits exactly the 8 tokens of one parity and the output is vector-aligned
                    const uint32_t *tv = dense0 ? hv[k] : ov[k];
                    if (wpos + 256 <= a.out_cap_tokens) stg_stream_v4(a.out + wpos + size_t(lane) * 8, make_uint4(tv[0], tv[1], tv[2], tv[3]));
                    else if (lane == 0) *a.scratch.overflow = 1u;
                    wpos += 256;
                    head = 0;
                    continue;
                }
                const uint32_t cnt = __popc(em);
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t tq = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl += tq;
                }
                const uint32_t total = __shfl_sync(FULL, incl, 31);
                uint32_t sp = stage_s + 2u * (pend + incl - cnt);  // where the lane's next token goes (before swizzling)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t v = (j & 1) ? ov[k][j >> 2] : hv[k][j >> 2];
                    const uint32_t tok = ((j >> 1) & 1) ? (v >> 16) : v;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                        "setp.ne.u32 p, %2, 0;\n\t"
                        "shr.u32 t, %0, 5;\n\t"
                        "and.b32 t, t, 0x7C;\n\t"
                        "xor.b32 t, t, %0;\n\t"
                        "@p st.shared.u16 [t], %1;\n\t"
                        "@p add.u32 %0, %0, 2;\n\t}"
                        : "+r"(sp)
                        : "h"(uint16_t(tok)), "r"(em & (1u << j))
                        : "memory");
                }
                __syncwarp();
                flush(total);
            }
            // the slice's last odd token (the next slice starts right behind it)
            if (pend > head && lane == 0) {
                uint32_t last;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(last) : "r"(stage_swz(stage_s)) : "memory");
                if (wpos + 1 <= a.out_cap_tokens) a.out[wpos] = uint16_t(last);
                else *a.scratch.overflow = 1u;
            }
        }
    }
}

template <int NG, int WG, int R>
struct FusedLaunch {
    using C = FusedCfg<NG, WG, R>;
    static cudaError_t configure(int dev) {
        static std::atomic<bool> configured[kMaxDevices];
        if (dev >= kMaxDevices) return cudaErrorInvalidDevice;
        if (!configured[dev].load(std::memory_order_acquire)) {
            cudaError_t err = cudaFuncSetAttribute(fused_sweep_kernel<NG, WG, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
            if (err != cudaSuccess) return err;
            configured[dev].store(true, std::memory_order_release);
        }
        return cudaSuccess;
    }
    // walls must lie on tile boundaries
    static bool applicable(const SweepArgs &a) {
        const size_t chunk = (a.chunk == 0 || a.chunk > a.n) ? a.n : a.chunk;
        return a.n != 0 && (chunk >= a.n || chunk % size_t(C::TILE) == 0);
    }
    static size_t n_tiles(size_t n) { return (n + C::TILE - 1) / C::TILE; }
    static cudaError_t launch(const SweepArgs &a, const uint16_t *d_table, cudaStream_t stream) {
        int dev = 0;
        cudaError_t err = cudaGetDevice(&dev);
        if (err != cudaSuccess) return err;
        err = configure(dev);
        if (err != cudaSuccess) return err;
        const size_t tiles = n_tiles(a.n);
        if (tiles * 8 > a.scratch.meta_bytes || tiles >= 0xfffffff0ull) return cudaErrorInvalidValue;
        err = cudaMemsetAsync(a.scratch.ctrl, 0, kCtrlBytes, stream);
        if (err != cudaSuccess) return err;
        err = cudaMemsetAsync(a.scratch.meta, 0, tiles * 8, stream);
        if (err != cudaSuccess) return err;
        size_t grid = (tiles + NG - 1) / NG;
        if (grid > size_t(sm_count(dev))) grid = size_t(sm_count(dev));
        const size_t chunk = (a.chunk == 0 || a.chunk > a.n) ? a.n : a.chunk;
        const uint32_t tpc = (chunk >= a.n) ? 0u : uint32_t(chunk / size_t(C::TILE));
        uint32_t *counter = reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(a.scratch.ctrl) + 320);
        fused_sweep_kernel<NG, WG, R><<<dim3(unsigned(grid)), dim3(C::THREADS), C::SMEM, stream>>>(
            a, d_table, reinterpret_cast<unsigned long long *>(a.scratch.meta), counter, uint32_t(tiles), tpc);
        return cudaGetLastError();
    }
};
