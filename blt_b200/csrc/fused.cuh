// fused.cuh -- the exact sweep in ONE pass: count, decoupled look-back and emit fused in a single persistent kernel.
//
// The input is read from DRAM once (N_in + 2*T_out bytes of traffic = the algorithmic bytes) and every pair is
// looked up once.  One CTA per SM holds the 128 KiB pair table and takes tiles of WG*R*512 input bytes in ticket
// order (one global counter).  WG worker warps and one chain warp (warp specialisation), one barrier per tile:
//
//   stage   the tile is copied into the CTA's shared-memory buffer by cp.async.bulk (issued by the chain warp as
//           soon as the workers have counted the previous tile, completion on an mbarrier);
//   count   every worker owns R*512 consecutive bytes of the tile (R rounds of 32 lanes x 16 bytes): both parities
//           are looked up, the tokens STAY IN REGISTERS, run parity is resolved inside the warp with two ballots
//           per round under the hypothesis "the warp's carry_in is 0" and the warp's slice is reduced to one carry
//           function (identity / constant, tokens for carry_in 0, the 0/1-token delta for carry_in 1);
//   chain   while the workers emit tile i-1 and count tile i+1, the chain warp composes the WG warp functions of
//           tile i, publishes the tile's function (status A) in the tile's 64-bit descriptor and polls the 64
//           descriptors in front of it until it sees an inclusive prefix (status P) with nothing missing behind it.
//           The carry entering every tile of the window comes from two ballots (nearest non-identity tile in front
//           of it), its exact token count is cnt0 - (delta & carry_in), and one warp-wide add gives the offset: the
//           chain advances up to 64 tiles per hop.  Chunk walls lie on tile boundaries, where the carry is 0 by
//           definition; chunk_ends fall out of the inclusive prefixes.  The look-back has a whole tile period of
//           slack before anybody needs its result, so nobody waits for it in the steady state;
//   emit    (one tile behind) every worker compacts its retained tokens into a warp-private staging line
//           (XOR-swizzled so that the 32 lanes' 2-byte stores spread over the banks) and streams whole 4-byte
//           words out.  Only the lanes in front of the slice's first non-identity segment depend on the carry_in;
//           they are redone when it is 1.
//
// Forward progress: a tile only ever waits for tiles with smaller tickets, which are held by resident CTAs, and
// nothing that publishes a tile's function waits for anything.
// Included by kernels.cu inside its anonymous namespace, after sweep3.cuh (ScanFn, scan_compose, start_bits).
#pragma once

constexpr unsigned long long FZ_A = 1ull << 62, FZ_P = 2ull << 62;
constexpr unsigned long long FZ_A_ID = 1ull << 61, FZ_A_CST = 1ull << 60, FZ_A_DELTA = 1ull << 59;
constexpr unsigned long long FZ_P_CARRY = 1ull << 60;  // same bit as FZ_A_CST: "the carry leaving this tile"
constexpr unsigned long long FZ_COUNT = (1ull << 56) - 1;
constexpr uint32_t FZ_F_START = 2u, FZ_NO_WALL = 0xffffffffu;

struct FusedShared {
    unsigned long long mbar;           // completion of the bulk copy into the tile buffer
    uint32_t tile[2];                  // ticket of iteration i in tile[i & 1]
    uint32_t flags[2];                 // FZ_F_START: it starts a chunk (the carry entering it is 0)
    uint32_t wall[2][2];               // offsets of the (at most two) chunk-last elements inside the tile, else FZ_NO_WALL:
                                       //   [0] a chunk boundary, [1] the end of the input if it is another element
    uint32_t len[2];                   // valid bytes of the tile (the tile size but for the input's last tile)
    unsigned long long wall_ck[2][2];  // the chunks those walls end
    unsigned long long fn_cnt[2][16];  // per worker: tokens of its slice for carry_in 0
    uint32_t fn_flags[2][16];          // per worker: bit0 identity, bit1 constant carry_out, bit2 delta
    unsigned long long res[2][16];     // per worker: carry_in << 63 | tokens of the launch in front of its slice
};

template <int WG, int R>
struct FusedCfg {
    static_assert(WG <= 16, "the chain warp scans the worker functions in one half warp");
    static_assert((WG * R * 512) % 16 == 0, "tiles start on 16-byte boundaries");
    static constexpr int THREADS = (WG + 1) * 32;  // WG workers + the chain warp
    static constexpr int WARP_BYTES = R * 512;
    static constexpr int TILE = WG * WARP_BYTES;
    static constexpr int BUF = TILE + 128;     // + the look-ahead vector
    static constexpr int STAGE_BYTES = 2048;   // per worker: 1 pending + 512 new tokens in a 1 KiB-aligned line
    static constexpr int OFF_STAGE = PairsFE::TABLE_BYTES;  // + up to 1 KiB of alignment slack
    static constexpr int OFF_BUF = OFF_STAGE + WG * STAGE_BYTES + 1024;
    static constexpr int OFF_GS = OFF_BUF + BUF;
    static constexpr int SMEM = OFF_GS + 1024;
    static_assert(sizeof(FusedShared) <= 1024, "control block");
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
__device__ __forceinline__ void cta_bar(int threads) { asm volatile("bar.sync 0, %0;" ::"r"(threads) : "memory"); }
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
// table read: the table is immutable once the CTA has passed its first barrier, so this asm is neither volatile nor
// a memory clobber (the compiler may schedule the 16 reads of a segment freely)
__device__ __forceinline__ uint32_t lds_tbl(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// staging line: shared address of a token -> where it really lives.  Bank bits 2-4 are XORed with the index of the
// 128-byte window (mod 8), so that stores 4 to 8 words apart (one lane's tokens behind the other's) do not pile up on
// a few banks.  Lines are 1 KiB aligned, so window i of a line is XORed with i & 7.
__device__ __forceinline__ uint32_t stage_swz(uint32_t addr) { return addr ^ ((addr >> 5) & 0x1Cu); }

// The SEG/2 pairs of one 16-byte segment that start at positions of parity PAR: the big-endian u16 to emit at each of
// those positions (merged id if the pair is a rule, else the element itself), two per register in position order.
template <int PAR>
__device__ __forceinline__ void fz_lookup(uint32_t tbl_s, const uint4 &w, uint32_t next, uint32_t *vals) {
    const uint32_t W[4] = {PAR ? __funnelshift_r(w.x, w.y, 8) : w.x, PAR ? __funnelshift_r(w.y, w.z, 8) : w.y,
                           PAR ? __funnelshift_r(w.z, w.w, 8) : w.z, PAR ? __funnelshift_r(w.w, next, 8) : w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t y = W[k] ^ ((W[k] >> 7) & 0x01FF01FFu);  // pair_table_index of both halves at once
        const uint32_t e0 = lds_tbl(tbl_s + ((y & 0xFFFFu) << 1));
        const uint32_t e1 = lds_tbl(tbl_s + ((y >> 16) << 1));
        vals[k] = __byte_perm(e0, e1, 0x5410);
    }
}

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

// Chain lane 0: starts the copy of tile `t` into the buffer.  Whole 16-byte vectors go through the bulk copy (tile +
// look-ahead vector where the input has them); the < 16 ragged bytes of the input's end are left to fz_copy_tail,
// which the whole chain warp runs.
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
// Where tile t meets chunk walls - chain lane 0.  Chunks are at least a tile long, so a tile holds at most one chunk
// boundary; the input's last tile may hold the end of the input as well.
template <class C>
__device__ __forceinline__ void fz_tile_geometry(const SweepArgs &a, unsigned long long chunk, uint32_t t, uint32_t n_tiles,
                                                 FusedShared *gs, uint32_t slot) {
    uint32_t flags = 0, wall0 = FZ_NO_WALL, wall1 = FZ_NO_WALL, len = 0;
    unsigned long long ck = 0;
    if (t < n_tiles) {
        const unsigned long long base = (unsigned long long)t * C::TILE;
        len = (a.n - base < (unsigned long long)C::TILE) ? uint32_t(a.n - base) : uint32_t(C::TILE);
        ck = base / chunk;
        if (base - ck * chunk == 0) flags |= FZ_F_START;
        const unsigned long long last = (ck + 1) * chunk - 1;  // the last element of the chunk the tile starts in
        if (last - base < (unsigned long long)len) wall0 = uint32_t(last - base);
        if (t + 1 == n_tiles && wall0 != len - 1u) wall1 = len - 1u;  // the input's last element ends the last chunk
    }
    gs->flags[slot] = flags;
    gs->len[slot] = len;
    gs->wall[slot][0] = wall0;
    gs->wall[slot][1] = wall1;
    gs->wall_ck[slot][0] = ck;
    gs->wall_ck[slot][1] = (a.n - 1) / chunk;
}

// ---- the chain warp: one call per tile -------------------------------------------------------------------------
// Composes the worker functions of tile `cur` (gs->fn_*[par]), publishes A, then resolves the tile's prefix by
// look-back, publishes P and leaves every worker's (carry_in, offset) in gs->res[par].
template <int WG>
__device__ __forceinline__ void fz_chain_tile(const SweepArgs &a, FusedShared *gs, unsigned long long *desc, uint32_t cur,
                                              uint32_t flags, uint32_t par, uint32_t n_tiles, int lane) {
    ScanFn item;
    item.id = 1; item.cst = 0; item.delta = 0; item.cnt0 = 0;
    if (lane < WG) {
        const uint32_t fl = gs->fn_flags[par][lane];
        item.id = fl & 1u; item.cst = (fl >> 1) & 1u; item.delta = (fl >> 2) & 1u;
        item.cnt0 = gs->fn_cnt[par][lane];
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
    if (lane == 0)
        st_desc(desc + cur, FZ_A | (tf.id ? FZ_A_ID : 0ull) | (tf.cst ? FZ_A_CST : 0ull) | (tf.delta ? FZ_A_DELTA : 0ull) | tf.cnt0);
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
            __nanosleep(64);
        }
    }
    if (starts) c_in = 0;
    const uint32_t c_out = tf.id ? c_in : tf.cst;
    const unsigned long long total = base + tf.cnt0 - ((c_in && tf.delta) ? 1ull : 0ull);
    if (lane == 0) {
        st_desc(desc + cur, FZ_P | (c_out ? FZ_P_CARRY : 0ull) | total);
        if (cur == n_tiles - 1) {
            *a.scratch.total_tokens = total;
            *a.scratch.merged_any = (total < a.n) ? 1u : 0u;
            if (a.out_base_tokens + total > a.out_cap_tokens) *a.scratch.overflow = 1u;
        }
    }
    if (lane < WG) {
        // (the scan ran on the worker functions as they are: a chunk start only fixes the carry entering worker 0)
        const uint32_t cw = ex.id ? c_in : ex.cst;
        const unsigned long long bw = base + ex.cnt0 - ((c_in && ex.delta) ? 1ull : 0ull);
        gs->res[par][lane] = (cw ? R_CARRY : 0ull) | bw;
    }
}

template <int WG, int R>
__global__ void __launch_bounds__((WG + 1) * 32, 1)
fused_sweep_kernel(const SweepArgs a, const uint16_t *__restrict__ table, unsigned long long *__restrict__ desc,
                   uint32_t *__restrict__ tile_counter, uint32_t n_tiles, unsigned long long chunk) {
    using C = FusedCfg<WG, R>;
    extern __shared__ __align__(16) unsigned char smem[];
    {  // the table, by everybody
        const uint4 *src = reinterpret_cast<const uint4 *>(table);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < PairsFE::TABLE_BYTES / 16; i += C::THREADS) dst[i] = src[i];
    }
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t smem_s = smem_u32(smem);
    const uint32_t tbl_s = smem_s;
    const uint32_t stage0_s = (smem_s + C::OFF_STAGE + 1023u) & ~1023u;  // 1 KiB-aligned staging lines
    FusedShared *gs = reinterpret_cast<FusedShared *>(smem + C::OFF_GS);
    unsigned char *buf = smem + C::OFF_BUF;
    const uint32_t bar = smem_u32(&gs->mbar);

    // ---- prologue: barrier, the first ticket and its copy ----------------------------------------------------
    uint32_t t_next = 0xffffffffu;  // chain lane 0: the ticket of the next iteration
    if (warp == WG) {
        if (lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        uint32_t first = 0;
        if (lane == 0) {
            first = atomicAdd(tile_counter, 1u);
            gs->tile[0] = first;
            fz_tile_geometry<C>(a, chunk, first, n_tiles, gs, 0);
            if (first < n_tiles) fz_issue_copy<C>(a, first, buf, bar);
            t_next = (first < n_tiles) ? atomicAdd(tile_counter, 1u) : first;
            gs->tile[1] = t_next;
            fz_tile_geometry<C>(a, chunk, t_next, n_tiles, gs, 1);
        }
        first = __shfl_sync(FULL, first, 0);
        if (first < n_tiles) fz_copy_tail<C>(a, first, buf, lane);
    }
    __syncthreads();  // table, barrier, first tickets

    if (warp == WG) {
        // =========================== chain warp ===========================
        uint32_t cur = gs->tile[0], flags = gs->flags[0];
        t_next = __shfl_sync(FULL, t_next, 0);
        bool prev_new = false;
        for (uint32_t it = 0;; ++it) {
            const uint32_t par = it & 1u;
            const bool have_new = cur < n_tiles;
            if (!have_new && !prev_new) break;  // (the workers leave on the same condition: equal barrier counts)
            cta_bar(C::THREADS);  // the workers have counted `cur`; they emit the tile before it now
            prev_new = have_new;
            if (!have_new) continue;
            // the buffer is free: the next tile may land in it (ragged tail first: the barrier's arrive releases it)
            if (t_next < n_tiles) {
                fz_copy_tail<C>(a, t_next, buf, lane);
                __syncwarp();
                if (lane == 0) fz_issue_copy<C>(a, t_next, buf, bar);
            }
            fz_chain_tile<WG>(a, gs, desc, cur, flags, par, n_tiles, lane);
            // ticket of the iteration after the next one, published before the next barrier
            cur = t_next;
            flags = gs->flags[par ^ 1u];
            uint32_t t2 = cur;
            if (lane == 0) {
                if (cur < n_tiles) t2 = atomicAdd(tile_counter, 1u);
                gs->tile[par] = t2;
                fz_tile_geometry<C>(a, chunk, t2, n_tiles, gs, par);
            }
            t_next = __shfl_sync(FULL, t2, 0);
        }
        return;
    }

    // =========================== workers ===========================
    const int wg = warp;
    const uint32_t stage_s = stage0_s + uint32_t(wg) * C::STAGE_BYTES;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t lane4 = uint32_t(lane) << 2;
    const unsigned char *slice = buf + wg * C::WARP_BYTES;
    const uint32_t slice_off = uint32_t(wg * C::WARP_BYTES);
    uint32_t parity = 0;
    // the tile counted in the previous iteration: tokens and emit masks stay in registers until its prefix is known
    uint32_t hvP[R][4], ovP[R][4], emP[R];
    bool prev_valid = false, prev_full = true;
    uint32_t prev_len = 0, prev_wall0 = FZ_NO_WALL, prev_wall1 = FZ_NO_WALL;
    unsigned long long prev_wall_ck0 = 0, prev_wall_ck1 = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        emP[k] = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) { hvP[k][q] = 0; ovP[k][q] = 0; }
    }

    for (uint32_t it = 0;; ++it) {
        const uint32_t par = it & 1u;
        const uint32_t cur = gs->tile[par];
        const bool have_new = cur < n_tiles;
        if (!have_new && !prev_valid) break;
        uint32_t hvN[R][4], ovN[R][4], emN[R];
        const uint32_t tile_len = gs->len[par];
        const uint32_t wall0 = gs->wall[par][0], wall1 = gs->wall[par][1];  // chunk-last elements in this tile, if any
        const unsigned long long wall_ck0 = gs->wall_ck[par][0], wall_ck1 = gs->wall_ck[par][1];
        const bool full = tile_len == uint32_t(C::TILE);
        if (have_new) {
            if (!mbar_wait(bar, parity)) *a.scratch.overflow = 3u;
            parity ^= 1u;
            // ---- count: lookups (retained), run parity under carry_in = 0, the slice's carry function ----
            bool t_id = true;
            uint32_t t_const = 0, delta = 0, cnt0 = 0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const uint32_t off = slice_off + uint32_t(k * 512 + lane * 16);  // offset of the lane's segment in the tile
                const uint4 w = *reinterpret_cast<const uint4 *>(slice + k * 512 + lane * 16);
                uint32_t next = __shfl_down_sync(FULL, w.x & 0xffu, 1);
                if (lane == 31) next = slice[k * 512 + 512];
                fz_lookup<0>(tbl_s, w, next, hvN[k]);
                fz_lookup<1>(tbl_s, w, next, ovN[k]);
                uint32_t valid = 0xFFFFu;
                if (!full) valid = (off + 16 <= tile_len) ? 0xFFFFu : (off < tile_len ? ((1u << (tile_len - off)) - 1u) : 0u);
                const uint32_t round_off = slice_off + uint32_t(k * 512);
                if (wall0 - round_off < 512u || wall1 - round_off < 512u) {  // warp-uniform: a wall is in this round
                    // a wall suppresses the pair that starts at the chunk's last element: the raw token goes out there
                    const uint32_t dj0 = wall0 - off, dj1 = wall1 - off;  // >= 16 (or wrapped) in every segment but one
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (dj0 == uint32_t(j) || dj1 == uint32_t(j)) {
                            const uint32_t be = PairsFE::raw_be(w, j);
                            uint32_t &dst = (j & 1) ? ovN[k][j >> 2] : hvN[k][j >> 2];
                            dst = ((j >> 1) & 1) ? ((dst & 0x0000ffffu) | (be << 16)) : ((dst & 0xffff0000u) | be);
                        }
                    }
                }
                const uint32_t m = fz_membership(hvN[k], ovN[k]) & valid;
                const uint32_t lead = __clz(~(m << 16));  // ones at the top of the segment
                const uint32_t nid = ~__ballot_sync(FULL, m == 0xFFFFu);
                const uint32_t cob = __ballot_sync(FULL, (lead & 1u) != 0);
                const uint32_t c_round0 = t_id ? 0u : t_const;
                const uint32_t l_nid = nid & lt_mask;
                const uint32_t cin0 = l_nid ? ((cob >> (31 - __clz(l_nid))) & 1u) : c_round0;
                const uint32_t st = start_bits(m, cin0);
                const uint32_t em = valid & ~((st << 1) | cin0);
                const uint32_t cnt = __popc(em);
                emN[k] = em;
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
                gs->fn_flags[par][wg] = (t_id ? 1u : 0u) | (t_const << 1) | (delta << 2);
                gs->fn_cnt[par][wg] = cnt0;
            }
        }
        cta_bar(C::THREADS);  // hands the tile to the chain warp; the previous tile's prefix is in gs->res[par ^ 1]

        // ---- emit (one tile behind): compaction of the retained tokens, streamed out in whole words -------------
        if (prev_valid) {
            const unsigned long long rv = gs->res[par ^ 1u][wg];
            const uint32_t slice_carry = uint32_t(rv >> 63);
            const unsigned long long abs0 = (rv & ~R_CARRY) + a.out_base_tokens;
            // stage[0 .. pend) holds tokens not yet written; logical token 0 of the line corresponds to a.out[wpos],
            // wpos is even.  `head`: that slot belongs to the slice in front of this one and is not written here.
            unsigned long long wpos = abs0 & ~1ull;
            uint32_t pend = uint32_t(abs0 & 1ull);
            uint32_t head = pend;
            unsigned long long rel = rv & ~R_CARRY;  // tokens of the launch in front of the next round
            bool dep = slice_carry != 0u;  // the lanes in front of the slice's first non-identity segment see carry_in = 1
            auto flush = [&](uint32_t total) {
                const uint32_t have = pend + total;
                const uint32_t nw = have >> 1;
                const bool fits = (wpos + have <= a.out_cap_tokens);
                if (!fits && lane == 0) *a.scratch.overflow = 1u;
                if (fits && nw != 0) {
                    unsigned char *gout = reinterpret_cast<unsigned char *>(a.out + wpos) + lane4;
                    {  // word v = lane + 32 i lives in window i of the line, XORed with i & 7
                        const uint32_t word = lds_u32(stage_s + lane4);
                        if (uint32_t(lane) < nw) {
                            if (lane == 0 && head != 0) *reinterpret_cast<uint16_t *>(gout + 2) = uint16_t(word >> 16);
                            else stg_stream_u32(gout, word);
                        }
                    }
#pragma unroll
                    for (int i = 1; i < 8; ++i) {
                        if (uint32_t(32 * i) < nw) {  // warp-uniform
                            const uint32_t word = lds_u32(stage_s + 128u * i + (lane4 ^ uint32_t((i & 7) << 2)));
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
                    if (odd && lane == 0) asm volatile("st.shared.u16 [%0], %1;" ::"r"(stage_s), "h"(uint16_t(keep)) : "memory");
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
                uint32_t em = emP[k];
                if (dep) {  // warp-uniform; false for good after the slice's first non-identity segment
                    const uint32_t off = slice_off + uint32_t(k * 512 + lane * 16);
                    uint32_t valid = 0xFFFFu;
                    if (!prev_full) valid = (off + 16 <= prev_len) ? 0xFFFFu : (off < prev_len ? ((1u << (prev_len - off)) - 1u) : 0u);
                    const uint32_t m = fz_membership(hvP[k], ovP[k]) & valid;
                    const uint32_t nid = ~__ballot_sync(FULL, m == 0xFFFFu);
                    if ((nid & lt_mask) == 0u) {
                        const uint32_t st1 = start_bits(m, 1u);
                        em = valid & ~((st1 << 1) | 1u);
                    }
                    if (nid) dep = false;
                }
                // a chunk that ends in this round: its output ends behind the token of the wall position
                const uint32_t round_off = slice_off + uint32_t(k * 512);
                const bool wall_here = (prev_wall0 - round_off < 512u || prev_wall1 - round_off < 512u) && a.chunk_ends != nullptr;  // warp-uniform
                const uint32_t dj0 = prev_wall0 - (round_off + uint32_t(lane * 16)), dj1 = prev_wall1 - (round_off + uint32_t(lane * 16));
                const uint32_t x = em ^ 0x5555u;
                const bool dense0 = __all_sync(FULL, x == 0u), dense1 = __all_sync(FULL, x == 0xFFFFu);
                if ((dense0 || dense1) && pend == 0 && (wpos & 7ull) == 0) {
                    // every lane emits exactly the 8 tokens of one parity and the output is vector-aligned
                    const uint32_t *tv = dense0 ? hvP[k] : ovP[k];
                    if (wpos + 256 <= a.out_cap_tokens) stg_stream_v4(a.out + wpos + size_t(lane) * 8, make_uint4(tv[0], tv[1], tv[2], tv[3]));
                    else if (lane == 0) *a.scratch.overflow = 1u;
                    if (wall_here) {
                        if (dj0 < 16u) a.chunk_ends[prev_wall_ck0] = a.chunk_ends_base + 2ull * (rel + 8u * lane + __popc(em & ((2u << dj0) - 1u)));
                        if (dj1 < 16u) a.chunk_ends[prev_wall_ck1] = a.chunk_ends_base + 2ull * (rel + 8u * lane + __popc(em & ((2u << dj1) - 1u)));
                    }
                    wpos += 256;
                    rel += 256;
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
                if (wall_here) {
                    if (dj0 < 16u) a.chunk_ends[prev_wall_ck0] = a.chunk_ends_base + 2ull * (rel + (incl - cnt) + __popc(em & ((2u << dj0) - 1u)));
                    if (dj1 < 16u) a.chunk_ends[prev_wall_ck1] = a.chunk_ends_base + 2ull * (rel + (incl - cnt) + __popc(em & ((2u << dj1) - 1u)));
                }
                rel += total;
                uint32_t sp = stage_s + 2u * (pend + incl - cnt);  // where the lane's next token goes (before swizzling)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t v = (j & 1) ? ovP[k][j >> 2] : hvP[k][j >> 2];
                    const uint32_t tok = ((j >> 1) & 1) ? (v >> 16) : v;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                        "setp.ne.u32 p, %2, 0;\n\t"
                        "shr.u32 t, %0, 5;\n\t"
                        "and.b32 t, t, 0x1C;\n\t"
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
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(last) : "r"(stage_s) : "memory");
                if (wpos + 1 <= a.out_cap_tokens) a.out[wpos] = uint16_t(last);
                else *a.scratch.overflow = 1u;
            }
        }
        // the tile just counted becomes the one to emit
        prev_valid = have_new;
        prev_full = full;
        prev_len = tile_len;
        prev_wall0 = wall0;
        prev_wall1 = wall1;
        prev_wall_ck0 = wall_ck0;
        prev_wall_ck1 = wall_ck1;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            emP[k] = emN[k];
#pragma unroll
            for (int q = 0; q < 4; ++q) { hvP[k][q] = hvN[k][q]; ovP[k][q] = ovN[k][q]; }
        }
    }
}

template <int WG, int R>
struct FusedLaunch {
    using C = FusedCfg<WG, R>;
    static cudaError_t configure(int dev) {
        static std::atomic<bool> configured[kMaxDevices];
        if (dev >= kMaxDevices) return cudaErrorInvalidDevice;
        if (!configured[dev].load(std::memory_order_acquire)) {
            cudaError_t err = cudaFuncSetAttribute(fused_sweep_kernel<WG, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
            if (err != cudaSuccess) return err;
            configured[dev].store(true, std::memory_order_release);
        }
        return cudaSuccess;
    }
    // at most one wall per tile; a tile starts on a 16-byte boundary of the input
    static bool applicable(const SweepArgs &a) {
        const size_t chunk = (a.chunk == 0 || a.chunk > a.n) ? a.n : a.chunk;
        return a.n != 0 && (chunk >= a.n || chunk >= size_t(C::TILE));
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
        size_t grid = tiles;
        if (grid > size_t(sm_count(dev))) grid = size_t(sm_count(dev));
        const size_t chunk = (a.chunk == 0 || a.chunk > a.n) ? a.n : a.chunk;
        uint32_t *counter = reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(a.scratch.ctrl) + 320);
        fused_sweep_kernel<WG, R><<<dim3(unsigned(grid)), dim3(C::THREADS), C::SMEM, stream>>>(
            a, d_table, reinterpret_cast<unsigned long long *>(a.scratch.meta), counter, uint32_t(tiles), (unsigned long long)chunk);
        return cudaGetLastError();
    }
};
