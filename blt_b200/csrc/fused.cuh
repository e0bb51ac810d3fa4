// fused.cuh -- the exact sweep in ONE pass: count, decoupled look-back and emit fused in a single persistent kernel.
//
// The input is read from DRAM once (N_in + 2*T_out bytes of traffic = the algorithmic bytes) and every pair is
// looked up once.  One CTA per SM holds the 128 KiB pair table; tiles of WG*R*512 input bytes are dealt to the CTAs
// round-robin.  WG worker warps and one chain warp (warp specialisation); they meet on mbarriers only (a worker
// never waits for another worker):
//
//   stage   every worker copies its own slice of the next tile into its shared-memory buffer with cp.async.bulk
//           (one elected lane, completion on the worker's mbarrier) as soon as it has counted the current one, so the
//           copy has the whole emit phase to land;
//   count   every worker owns R*512 consecutive bytes of the tile (R rounds of 32 lanes x 16 bytes): both parities
//           are looked up, the tokens STAY IN REGISTERS, run parity is resolved inside the warp with two ballots
//           per round under the hypothesis "the warp's carry_in is 0" and the warp's slice is reduced to one carry
//           function (identity / constant, tokens for carry_in 0, the 0/1-token delta for carry_in 1).  The last
//           worker to finish composes the WG functions and publishes the tile's function (status A) in the tile's
//           64-bit descriptor.  (Looking all R rounds up first and doing the ballot algebra afterwards was measured:
//           no gain, 1.32 against 1.30 ms per GiB on config 2.)
//   chain   while the workers emit tile i-1 and count tile i+1, the chain warp polls the 192 descriptors in front of
//           tile i until it sees an inclusive prefix (status P) with nothing missing behind it.  The carry entering
//           every tile of the window comes from ballots (nearest non-identity tile in front of it), its exact token
//           count is cnt0 - (delta & carry_in), and one warp-wide add gives the offset.  The window spans more than
//           one round of the deal, so nothing propagates hop by hop inside a round;
//   emit    (one tile behind) every worker compacts the retained tokens of its whole slice into a warp-private
//           staging line with 2-byte stores predicated by the emit mask (the R rounds are R independent store chains)
//           and sends the whole 16-byte vectors out with one bulk copy shared -> global.  Only
//           the lanes in front of the slice's first non-identity segment depend on the carry_in; they are redone when
//           it is 1.  A slice that emits exactly one parity everywhere goes out straight from the registers.
//
// Forward progress: a tile only ever waits for tiles with smaller numbers, which belong to resident CTAs that reach
// them before any larger one, and nothing that publishes a tile's function waits for anything.
// Measurements, the per-round budget and the variants that were tried: DESIGN.md section 4.
// Included by kernels.cu inside its anonymous namespace, after sweep3.cuh (ScanFn, scan_compose, start_bits).
#pragma once

constexpr unsigned long long FZ_A = 1ull << 62, FZ_P = 2ull << 62;
constexpr unsigned long long FZ_A_ID = 1ull << 61, FZ_A_CST = 1ull << 60, FZ_A_DELTA = 1ull << 59;
constexpr unsigned long long FZ_P_CARRY = 1ull << 60;  // same bit as FZ_A_CST: "the carry leaving this tile"
constexpr unsigned long long FZ_COUNT = (1ull << 56) - 1;
constexpr uint32_t FZ_F_START = 2u, FZ_NO_WALL = 0xffffffffu;

// BLT_FUSED_PROF builds only: per worker (CTA x 16 + worker) x 8 counters, in clock cycles:
// 0 waiting for the slice copy, 1 counting, 2 waiting for the chain warp, 3 emitting, 4 tiles, 5 SM id
#ifdef BLT_FUSED_PROF
__device__ unsigned long long g_fz_prof[256 * 32 * 8];
#define FZ_PROF(stmt) stmt
#else
#define FZ_PROF(stmt)
#endif

constexpr int FZ_DMAX = 4;  // deepest pipeline: tiles a worker holds between "counted" and "emitted"

struct FusedShared {
    unsigned long long wbar[32];              // per worker: completion of the bulk copy of its slice
    // per tile in flight, slot = iteration % D:
    unsigned long long counted[FZ_DMAX];      // the tile is counted and its function published: one arrival
    unsigned long long resolved[FZ_DMAX];     // the chain warp has left that tile's prefix in res[slot]: one arrival
    uint32_t arrived[FZ_DMAX];                // workers that have counted it (the last one composes and publishes the function)
    uint32_t tf_flags[FZ_DMAX];               // the tile's function, left for the chain warp by that worker (+ bit3: starts a chunk)
    unsigned long long tf_cnt[FZ_DMAX];
    uint32_t ex_flags[FZ_DMAX][32];           // per worker: the workers in front of it composed
    unsigned long long ex_cnt[FZ_DMAX][32];
    unsigned long long fn_cnt[FZ_DMAX][32];   // per worker: tokens of its slice for carry_in 0
    uint32_t fn_flags[FZ_DMAX][32];           // per worker: bit0 identity, bit1 constant carry_out, bit2 delta
    unsigned long long res[FZ_DMAX][32];      // per worker: carry_in << 63 | tokens of the launch in front of its slice
    // geometry, slot = iteration % (4 D): written D+1 iterations ahead (when the tile is claimed), read when the worker
    // starts the copy of its slice, when the tile is counted and when it is emitted
    uint32_t tile_id[4 * FZ_DMAX];            // the tile of that iteration (claimed from the launch's counter); >= n_tiles: none
    uint32_t flags[4 * FZ_DMAX];              // FZ_F_START: it starts a chunk (the carry entering it is 0)
    uint32_t len[4 * FZ_DMAX];                // valid bytes of the tile (the tile size but for the input's last tile)
    uint32_t wall[4 * FZ_DMAX][2];            // offsets of the (at most two) chunk-last elements inside the tile, else FZ_NO_WALL:
                                              //   [0] a chunk boundary, [1] the end of the input if it is another element
    unsigned long long wall_ck[4 * FZ_DMAX][2];  // the chunks those walls end
    FZ_PROF(long long t_first[FZ_DMAX]; long long t_pub[FZ_DMAX];)  // first / last worker done with the tile (clock64)
};

template <int WG, int R, int D>
struct FusedCfg {
    static_assert(D == 2 || D == 4, "pipeline depth (tiles per worker between counted and emitted + 1)");
    static_assert(WG <= 31, "the chain warp scans the worker functions in one warp");
    static_assert((WG * R * 512) % 16 == 0, "tiles start on 16-byte boundaries");
    static_assert(R % 2 == 0, "the emit scan packs two rounds per word");
    static constexpr int THREADS = (WG + 1) * 32;  // WG workers + the chain warp
    static constexpr int LB = 6;  // look-back window: 32 * LB tiles (more than the 148 tiles of one round of the deal)
    static constexpr int WARP_BYTES = R * 512;
    static constexpr int TILE = WG * WARP_BYTES;
    static constexpr int WBUF = WARP_BYTES + 32;  // a worker's slice + the look-ahead vector
    static constexpr int BUF = WG * WBUF;
    static constexpr int STAGE_BYTES = (2 * (8 + R * 512) + 127) / 128 * 128;  // per worker: up to 7 head slots + R*512 tokens + 1 scratch slot
    static constexpr int OFF_STAGE = PairsFE::TABLE_BYTES;  // + up to 128 bytes of alignment slack
    static constexpr int OFF_BUF = OFF_STAGE + WG * STAGE_BYTES + 128;
    static constexpr int OFF_GS = OFF_BUF + BUF;
    static constexpr int SMEM = OFF_GS + int((sizeof(FusedShared) + 127) / 128 * 128);
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
// false: the phase did not complete within seconds (reported as a CUDA error by the host instead of hanging the device).
// The warp is suspended by the hardware between two looks (try_wait with a suspend-time hint): a waiting warp issues
// one instruction per wake-up instead of spinning through the issue slots and the alu pipe of the working ones.
template <int HINT_NS>
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t tries = 0; tries < (1u << 22); ++tries) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(uint32_t(HINT_NS))
            : "memory");
        if (ok != 0u) return true;
    }
    return false;
}
// global -> shared bulk copy (the TMA unit's 1-D form): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long *p, unsigned long long v) {
#ifdef BLT_FZ_ATOM
    asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %0;" : "+l"(v) : "l"(p) : "memory");
#else
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
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
// staging line: shared address of a token -> where it really lives.  The index of the 16-byte vector inside its
// 128-byte window (address bits 4-6) is XORed with the index of the window (mod 8), so that the lanes' 2-byte stores
// (one lane's tokens about 17 bytes behind the other's) do not pile up on a few banks while 16-byte vectors stay whole
// for the flush.  Every 128-byte window is permuted within itself: lines are whole, 128-byte aligned windows.
#ifdef BLT_FZ_STAGE_SWIZZLE
constexpr bool kStageSwizzle = true;
#else
constexpr bool kStageSwizzle = false;  // measured: the swizzle costs 2 instructions per position and does not pay (config 2: 1.39 -> 1.30 ms per GiB without it)
#endif
__device__ __forceinline__ uint32_t stage_swz(uint32_t addr) { return kStageSwizzle ? (addr ^ ((addr >> 3) & 0x70u)) : addr; }
#ifndef BLT_FZ_NO_BULK_FLUSH
static_assert(!kStageSwizzle, "the bulk-copy flush reads the staging line as it lies");
#endif

// The SEG/2 pairs of one 16-byte segment that start at positions of parity PAR: the big-endian u16 to emit at each of
// those positions (merged id if the pair is a rule, else the element itself), two per register in position order.
// (Integer shifts and adds go to the alu pipe, which takes one warp instruction every two cycles per scheduler and is
// this kernel's busiest unit; multiply-high by a power of two is a right shift on the fma pipe, multiply-add an address
// computation there.)
__device__ __forceinline__ uint32_t shr_fma(uint32_t x, uint32_t two_pow_32_minus_k) {
    uint32_t r;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(two_pow_32_minus_k));
    return r;
}
template <int PAR>
__device__ __forceinline__ void fz_lookup(uint32_t tbl_s, const uint4 &w, uint32_t next, uint32_t *vals) {
    const uint32_t W[4] = {PAR ? __funnelshift_r(w.x, w.y, 8) : w.x, PAR ? __funnelshift_r(w.y, w.z, 8) : w.y,
                           PAR ? __funnelshift_r(w.z, w.w, 8) : w.z, PAR ? __funnelshift_r(w.w, next, 8) : w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t y = W[k] ^ (shr_fma(W[k], 1u << 25) & 0x01FF01FFu);  // pair_table_index of both halves at once
        // entry address = table + 2 * index: one multiply-add per entry (the compiler's own form is shift, mask, add)
        uint32_t a0, a1;
        asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(a0) : "r"(y & 0xFFFFu), "r"(tbl_s));
        asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(a1) : "r"(shr_fma(y, 1u << 16)), "r"(tbl_s));
        vals[k] = __byte_perm(lds_tbl(a0), lds_tbl(a1), 0x5410);
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

// A worker starts the copy of its slice of tile `t` into its own buffer.  Whole 16-byte vectors go through the bulk
// copy (slice + look-ahead vector where the input has them); the < 16 ragged bytes at the input's end are copied by
// the lanes themselves first (the barrier's arrive releases them).
template <class C>
__device__ __forceinline__ void fz_warp_copy(const SweepArgs &a, uint32_t t, int wg, unsigned char *wbuf, uint32_t wbar, int lane) {
    const unsigned long long base = (unsigned long long)t * C::TILE + (unsigned long long)wg * C::WARP_BYTES;
    uint32_t avail = 0;
    if (base < a.n) {
        const unsigned long long left = a.n - base;
        avail = left < (unsigned long long)(C::WARP_BYTES + 16) ? uint32_t(left) : uint32_t(C::WARP_BYTES + 16);
    }
    const uint32_t bytes16 = avail & ~15u;
    if (bytes16 + lane < avail) wbuf[bytes16 + lane] = static_cast<const unsigned char *>(a.in)[base + bytes16 + lane];
    __syncwarp();
    if (lane == 0) {
        if (bytes16 != 0) {
            mbar_expect_tx(wbar, bytes16);
            bulk_g2s(smem_u32(wbuf), static_cast<const unsigned char *>(a.in) + base, bytes16, wbar);
        } else {
            mbar_arrive(wbar);
        }
    }
}

// Where tile t meets chunk walls - chain lane 0.  Chunks are at least a tile long, so a tile holds at most one chunk
// boundary; the input's last tile may hold the end of the input as well.
// ck_hint / rem_hint: the chunk the tile starts in and its offset there when the caller tracks them incrementally
// (rem_hint = ~0: not known, divide).
template <class C>
__device__ __forceinline__ void fz_tile_geometry(const SweepArgs &a, unsigned long long chunk, unsigned long long t, uint32_t n_tiles,
                                                 unsigned long long last_ck, FusedShared *gs, uint32_t slot,
                                                 unsigned long long ck_hint = 0, unsigned long long rem_hint = ~0ull) {
    uint32_t flags = 0, wall0 = FZ_NO_WALL, wall1 = FZ_NO_WALL, len = 0;
    unsigned long long ck = 0;
    if (t < n_tiles) {
        const unsigned long long base = t * C::TILE;
        ck = (rem_hint != ~0ull) ? ck_hint : base / chunk;
        const unsigned long long rem = (rem_hint != ~0ull) ? rem_hint : base - ck * chunk;  // the tile's offset in the chunk it starts in
        len = (a.n - base < (unsigned long long)C::TILE) ? uint32_t(a.n - base) : uint32_t(C::TILE);
        if (rem == 0) flags |= FZ_F_START;
        const unsigned long long to_last = chunk - 1 - rem;  // distance to the last element of the chunk the tile starts in
        if (to_last < (unsigned long long)len) wall0 = uint32_t(to_last);
        if (t + 1 == n_tiles && wall0 != len - 1u) wall1 = len - 1u;  // the input's last element ends the last chunk
    }
    gs->tile_id[slot] = t < n_tiles ? uint32_t(t) : 0xffffffffu;
    gs->flags[slot] = flags;
    gs->len[slot] = len;
    gs->wall[slot][0] = wall0;
    gs->wall[slot][1] = wall1;
    gs->wall_ck[slot][0] = ck;
    gs->wall_ck[slot][1] = last_ck;
}

// The tile of iteration `iter` of this CTA (chain lane 0).  Static deal (DYN = false, the default): tile blockIdx.x +
// iter * gridDim.x.  Dynamic (DYN = true): the next tile of one counter of the launch, so that a tile's predecessors
// were claimed before it whatever the phase of their CTAs - measured SLOWER (tiles are claimed D+1 iterations before
// they are counted, which loosens the order again: config 3 exact 1.00 -> 1.15 ms per GiB), kept for the record.
template <bool DYN>
__device__ __forceinline__ unsigned long long fz_claim(unsigned int *counter, uint32_t n_tiles, uint32_t iter) {
    const unsigned long long t = DYN ? (unsigned long long)atomicAdd(counter, 1u)
                                     : (unsigned long long)blockIdx.x + (unsigned long long)iter * gridDim.x;
    return t < n_tiles ? t : 0xffffffffull;
}

// ---- the chain warp ------------------------------------------------------------------------------------------
// A tile between "its function is published" and "its prefix is resolved".
struct FzPending {
    ScanFn tf;       // the whole tile
    ScanFn ex;       // lane w: workers 0 .. w-1 composed
    uint32_t cur, par, starts;
};

// Composes the worker functions of tile `cur` (gs->fn_*[par]) and publishes the tile's function (status A).  Never
// waits for anything: other tiles' look-backs depend on it.
template <int WG>
__device__ __forceinline__ FzPending fz_chain_publish(FusedShared *gs, unsigned long long *desc, uint32_t cur, uint32_t flags,
                                                      uint32_t par, int lane) {
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
    FzPending pd;
    pd.ex = scan_shfl_up(inc, 1);
    if (lane == 0) { pd.ex.id = 1; pd.ex.cst = 0; pd.ex.delta = 0; pd.ex.cnt0 = 0; }
    {
        const uint32_t packed = inc.id | (inc.cst << 1) | (inc.delta << 2);
        const uint32_t p = __shfl_sync(FULL, packed, WG - 1);
        pd.tf.id = p & 1u; pd.tf.cst = (p >> 1) & 1u; pd.tf.delta = (p >> 2) & 1u;
        pd.tf.cnt0 = __shfl_sync(FULL, inc.cnt0, WG - 1);
    }
    pd.cur = cur; pd.par = par;
    pd.starts = (flags & FZ_F_START) ? 1u : 0u;
    if (pd.starts) {  // the carry entering a chunk is 0: the tile's function collapses to a constant
        pd.tf.cst = pd.tf.id ? 0u : pd.tf.cst;
        pd.tf.id = 0; pd.tf.delta = 0;
    }
    if (lane == 0)
        st_desc(desc + cur, FZ_A | (pd.tf.id ? FZ_A_ID : 0ull) | (pd.tf.cst ? FZ_A_CST : 0ull) | (pd.tf.delta ? FZ_A_DELTA : 0ull) | pd.tf.cnt0);
    return pd;
}

// One poll of the look-back of a pending tile.  On success: publishes its inclusive prefix (status P), leaves every
// worker's (carry_in, offset) in gs->res[par] and returns true.
// Position x <-> tile cur-1-x; lane i holds positions i, i+32, ... (word j of a mask = positions 32j .. 32j+31).
template <int WG, int LB>
__device__ __forceinline__ bool fz_chain_poll(const SweepArgs &a, FusedShared *gs, unsigned long long *desc, const FzPending &pd,
                                              uint32_t n_tiles, int lane, uint32_t *prof_holes = nullptr, uint32_t *prof_q = nullptr) {
    const uint32_t cur = pd.cur;
    unsigned long long d[LB];
#pragma unroll
    for (int j = 0; j < LB; ++j) {
        const long long idx = (long long)cur - 1 - lane - 32 * j;
        d[j] = FZ_P;  // in front of tile 0: carry 0, nothing emitted
        if (idx >= 0) d[j] = ld_desc(desc + idx);
    }
    int qw = -1, qb = 0;  // word and bit of the nearest inclusive prefix
    // (every branch below is warp-uniform: ballots; a poll that cannot succeed ends after the word with the hole)
#pragma unroll
    for (int j = 0; j < LB; ++j) {
        if (qw < 0) {
            const uint32_t st = uint32_t(d[j] >> 62);
            const uint32_t pmj = __ballot_sync(FULL, st == 2u);
            const uint32_t zmj = __ballot_sync(FULL, st == 0u);
            if (pmj != 0u) {
                qw = j;
                qb = __ffs(pmj) - 1;
                if (zmj & ((1u << qb) - 1u)) {  // a tile nearer than the prefix has published nothing yet
                    FZ_PROF(if (prof_holes != nullptr) { *prof_holes = __popc(zmj & ((1u << qb) - 1u)); *prof_q = uint32_t(32 * qw + qb); })
                    return false;
                }
            } else if (zmj != 0u) {
                FZ_PROF(if (prof_holes != nullptr) { *prof_holes = __popc(zmj); *prof_q = 999u; })
                return false;
            }
        }
    }
    if (qw < 0) {
        FZ_PROF(if (prof_holes != nullptr) { *prof_holes = 0; *prof_q = 999u; })
        return false;
    }
    FZ_PROF(if (prof_holes != nullptr) { *prof_holes = 0; *prof_q = uint32_t(32 * qw + qb); })
    const int q = 32 * qw + qb;
    // The carry leaving position x: P.carry at q, the constant of a non-identity tile, else whatever enters it.
    // nim = "non-identity" (with q itself), cm = that carry.  up_c[j]: the carry of the nearest non-identity position
    // in words >= j (warp-uniform, filled from the far end).  Words behind the prefix's word take no part.
    uint32_t nim[LB], cm[LB];
#pragma unroll
    for (int j = 0; j < LB; ++j) {
        nim[j] = 0u; cm[j] = 0u;
        if (j <= qw) {
            nim[j] = __ballot_sync(FULL, (d[j] & FZ_A_ID) == 0ull);
            cm[j] = __ballot_sync(FULL, (d[j] & FZ_A_CST) != 0ull);
            if (j == qw) nim[j] = (nim[j] & ((1u << qb) - 1u)) | (1u << qb);
        }
    }
    uint32_t up_c[LB + 1];
    up_c[LB] = 0u;
#pragma unroll
    for (int j = LB - 1; j >= 0; --j) {
        up_c[j] = up_c[j + 1];
        if (nim[j] != 0u) up_c[j] = (cm[j] >> (__ffs(nim[j]) - 1)) & 1u;
    }
    uint32_t e = 0;
#pragma unroll
    for (int j = 0; j < LB; ++j) {
        if (j <= qw) {
            // the carry entering position 32j + lane leaves the nearest non-identity position behind it
            const uint32_t above = (lane < 31) ? (nim[j] >> (lane + 1)) : 0u;
            const uint32_t cin = above ? ((cm[j] >> (lane + __ffs(above))) & 1u) : up_c[j + 1];
            if (32 * j + lane < q) e += uint32_t(d[j] & 0xffffffffull) - (((d[j] & FZ_A_DELTA) && cin) ? 1u : 0u);
        }
    }
    e = __reduce_add_sync(FULL, e);
    unsigned long long pdesc = 0;
#pragma unroll
    for (int j = 0; j < LB; ++j)
        if (j == qw) pdesc = __shfl_sync(FULL, d[j], qb);
    const unsigned long long base = (pdesc & FZ_COUNT) + e;
    const uint32_t c_in = pd.starts ? 0u : up_c[0];  // the nearest non-identity position's carry enters this tile
    const uint32_t c_out = pd.tf.id ? c_in : pd.tf.cst;
    const unsigned long long total = base + pd.tf.cnt0 - ((c_in && pd.tf.delta) ? 1ull : 0ull);
    if (lane == 0) {
        st_desc(desc + cur, FZ_P | (c_out ? FZ_P_CARRY : 0ull) | total);
        if (cur == n_tiles - 1) {
            *a.scratch.total_tokens = total + a.total_bias;
            *a.scratch.merged_any = (total < a.n || a.total_bias != 0) ? 1u : 0u;
            if (a.out_base_tokens + total > a.out_cap_tokens) *a.scratch.overflow = 1u;
        }
    }
    if (lane < WG) {
        // (the scan ran on the worker functions as they are: a chunk start only fixes the carry entering worker 0)
        const uint32_t cw = pd.ex.id ? c_in : pd.ex.cst;
        const unsigned long long bw = base + pd.ex.cnt0 - ((c_in && pd.ex.delta) ? 1ull : 0ull);
        gs->res[pd.par][lane] = (cw ? R_CARRY : 0ull) | bw;
    }
    return true;
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0u;
}

template <int V>
struct FzIC { static constexpr int value = V; };

template <int WG, int R, int D>
__global__ void __launch_bounds__((WG + 1) * 32, 1)
fused_sweep_kernel(const SweepArgs a, const uint16_t *__restrict__ table, unsigned long long *__restrict__ desc,
                   uint32_t n_tiles, unsigned long long chunk) {
    using C = FusedCfg<WG, R, D>;
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
    const uint32_t stage0_s = (smem_s + C::OFF_STAGE + 127u) & ~127u;  // 128-byte aligned staging lines
    FusedShared *gs = reinterpret_cast<FusedShared *>(smem + C::OFF_GS);
    unsigned char *buf = smem + C::OFF_BUF;
    const uint32_t bar_counted = smem_u32(&gs->counted[0]), bar_resolved = smem_u32(&gs->resolved[0]);

    // ---- prologue: barriers, the first D+1 tiles -----------------------------------------------------------------
    // Tiles are dealt round-robin (iteration i of this CTA is tile blockIdx.x + i * gridDim.x); their geometry is laid
    // out D+1 iterations ahead of their count (the worker starts the copy of its next slice one iteration ahead, before
    // it has met the chain warp again).  Iteration i uses slot i % D of the per-tile state and slot i % 4D of the geometry.
    constexpr uint32_t GS = 4 * D;
#ifdef BLT_FZ_DYNAMIC_TILES
    constexpr bool DYN = true;
#else
    constexpr bool DYN = false;
#endif
    unsigned int *claim = reinterpret_cast<unsigned int *>(static_cast<unsigned char *>(a.scratch.ctrl) + 128);
    if (warp == WG) {
        if (lane == 0) {
            for (int w = 0; w < WG; ++w) mbar_init(smem_u32(&gs->wbar[w]), 1);
            for (int d = 0; d < D; ++d) {
                mbar_init(bar_counted + 8 * d, 1);
                mbar_init(bar_resolved + 8 * d, 1);
                gs->arrived[d] = 0u;
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const unsigned long long last_ck = (a.n - 1) / chunk;
            for (uint32_t j = 0; j <= uint32_t(D); ++j) fz_tile_geometry<C>(a, chunk, fz_claim<DYN>(claim, n_tiles, j), n_tiles, last_ck, gs, j);
        }
    }
    __syncthreads();  // table, barriers, first geometry: the only CTA-wide barrier of the kernel

    if (warp == WG) {
        // =========================== chain warp ===========================
        // Resolves the tiles whose functions are published, oldest first.  Up to D tiles are pending: the workers cannot
        // count tile i before they have emitted tile i-D.
        const unsigned long long last_ck = (a.n - 1) / chunk;
        FzPending pend[D];
#pragma unroll
        for (int d = 0; d < D; ++d) pend[d].cur = 0;
        uint32_t n_pend = 0;            // pend[0] is the oldest one
        uint32_t it = 0;                 // the next iteration of this CTA to be counted
        uint32_t idle = 0;
        // Geometry of the iterations ahead: iteration g_it (the next one to lay out) is tile g_t, which starts g_rem bytes
        // into chunk g_ck; the static deal advances by gridDim.x tiles per iteration, so no division per tile.
        uint32_t owed = 0, g_it = uint32_t(D) + 1u;
        const unsigned long long step_bytes = (unsigned long long)gridDim.x * C::TILE;
        const unsigned long long step_ck = step_bytes / chunk, step_rem = step_bytes % chunk;
        unsigned long long g_t = (unsigned long long)blockIdx.x + (unsigned long long)g_it * gridDim.x;
        unsigned long long g_ck = (g_t * C::TILE) / chunk, g_rem = (g_t * C::TILE) % chunk;
        auto lay_out = [&]() {
            while (owed != 0) {  // warp-uniform
                if (lane == 0) {
                    if (DYN) fz_tile_geometry<C>(a, chunk, fz_claim<true>(claim, n_tiles, g_it), n_tiles, last_ck, gs, g_it % GS);
                    else fz_tile_geometry<C>(a, chunk, g_t < n_tiles ? g_t : 0xffffffffull, n_tiles, last_ck, gs, g_it % GS, g_ck, g_rem);
                }
                g_t += gridDim.x;
                g_ck += step_ck;
                g_rem += step_rem;
                if (g_rem >= chunk) { g_rem -= chunk; ++g_ck; }
                ++g_it;
                --owed;
            }
            __syncwarp();
        };
        FZ_PROF(long long pc_spread = 0; long long pc_pick = 0; long long pc_res = 0; long long pc_polls = 0; long long pc_tiles = 0; long long t_pick[D];
                long long pc_poll_cyc = 0; long long pc_first_holes = 0; long long pc_first_q = 0; long long pc_first_age = 0; long long pc_firsts = 0;
                uint32_t pc_last_polled = 0xffffffffu;)
        while (true) {
            const uint32_t nxt = gs->tile_id[it % GS];  // its tile, if there is one
            if (nxt == 0xffffffffu && n_pend == 0) break;
            bool progress = false;
            if (nxt != 0xffffffffu && n_pend < uint32_t(D) && mbar_test(bar_counted + 8 * (it % D), (it / D) & 1u)) {
                // tile `nxt` is counted and its function published (by the last worker to finish): take it over
                const uint32_t slot = it % D;
                FzPending p;
                const uint32_t tfl = gs->tf_flags[slot];
                p.tf.id = tfl & 1u; p.tf.cst = (tfl >> 1) & 1u; p.tf.delta = (tfl >> 2) & 1u; p.starts = (tfl >> 3) & 1u;
                p.tf.cnt0 = gs->tf_cnt[slot];
                const uint32_t efl = gs->ex_flags[slot][lane];
                p.ex.id = efl & 1u; p.ex.cst = (efl >> 1) & 1u; p.ex.delta = (efl >> 2) & 1u;
                p.ex.cnt0 = gs->ex_cnt[slot][lane];
                p.cur = nxt; p.par = slot;
                FZ_PROF({ const long long now = clock64(); pc_spread += gs->t_pub[slot] - gs->t_first[slot]; pc_pick += now - gs->t_pub[slot];
                          for (int d = 0; d < D; ++d) if (uint32_t(d) == n_pend) t_pick[d] = gs->t_pub[slot]; ++pc_tiles; })
                // the geometry of iteration it+D+1 is laid out below, behind the poll (owed = iterations picked up whose
                // geometry is still to be written; the `resolved` arrive of the tile releases it to the workers)
                ++owed;
#pragma unroll
                for (int d = 0; d < D; ++d)
                    if (uint32_t(d) == n_pend) pend[d] = p;
                ++n_pend;
                ++it;
                progress = true;
            }
            FZ_PROF(if (n_pend != 0) ++pc_polls;)
            FZ_PROF(uint32_t ph = 0; uint32_t pq = 0; const long long tp0 = clock64(); const bool first_poll = n_pend != 0 && pend[0].cur != pc_last_polled;)
#ifdef BLT_FUSED_PROF
            const bool poll_ok = n_pend != 0 && fz_chain_poll<WG, C::LB>(a, gs, desc, pend[0], n_tiles, lane, &ph, &pq);
            if (n_pend != 0) {
                pc_poll_cyc += clock64() - tp0;
                if (first_poll) { pc_first_holes += ph; pc_first_q += pq; pc_first_age += tp0 - t_pick[0]; ++pc_firsts; pc_last_polled = pend[0].cur; }
            }
            if (poll_ok) {
#else
            if (n_pend != 0 && fz_chain_poll<WG, C::LB>(a, gs, desc, pend[0], n_tiles, lane)) {
#endif
                const uint32_t slot = pend[0].par;
                FZ_PROF({ pc_res += clock64() - t_pick[0]; for (int d = 0; d + 1 < D; ++d) t_pick[d] = t_pick[d + 1]; })
                lay_out();  // (only if this tile was picked up in this very turn of the loop)
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_resolved + 8 * slot);  // releases res[slot] and the geometry written so far
#pragma unroll
                for (int d = 0; d + 1 < D; ++d) pend[d] = pend[d + 1];
                --n_pend;
                progress = true;
            }
            lay_out();
            if (!progress) {
                if (++idle == (1u << 24)) {  // seconds: somebody died; fail the launch instead of hanging
                    *a.scratch.overflow = 3u;
                    break;
                }
                __nanosleep(20);  // (0 and 150 ns measured: the same time)
            } else {
                idle = 0;
            }
        }
        FZ_PROF(if (lane == 0 && blockIdx.x < 256) {
            unsigned long long *pp = g_fz_prof + (size_t(blockIdx.x) * 32 + 31) * 8;
            pp[0] = pc_spread; pp[1] = pc_pick; pp[2] = pc_res; pp[3] = pc_polls; pp[4] = pc_tiles;
            unsigned long long *pq2 = g_fz_prof + (size_t(blockIdx.x) * 32 + 30) * 8;
            pq2[0] = pc_poll_cyc; pq2[1] = pc_first_holes; pq2[2] = pc_first_q; pq2[3] = pc_first_age; pq2[4] = pc_firsts;
        })
        return;
    }

    // =========================== workers ===========================
    const int wg = warp;
    const uint32_t stage_s = stage0_s + uint32_t(wg) * C::STAGE_BYTES;
    const uint32_t lt_mask = (1u << lane) - 1u;
    unsigned char *slice = buf + wg * C::WBUF;  // this worker's slice of the tile being counted
    const uint32_t wbar = smem_u32(&gs->wbar[wg]);
    const uint32_t slice_off = uint32_t(wg * C::WARP_BYTES);
    uint32_t parity = 0;
    if (gs->tile_id[0] != 0xffffffffu) fz_warp_copy<C>(a, gs->tile_id[0], wg, slice, wbar, lane);
    // D sets of retained tiles: tokens and emit masks stay in registers until the tile's prefix is known.  Iteration
    // `it` counts into set it % D and emits the tile of iteration it - (D-1) from set (it + 1) % D.
    uint32_t hv[D][R][4], ov[D][R][4], em[D][R];
    FZ_PROF(long long pf_copy = 0; long long pf_count = 0; long long pf_chain = 0; long long pf_emit = 0; long long pf_tiles = 0; long long pf_t = clock64();)

    auto iteration = [&](auto ns_c, auto os_c, uint32_t it) -> bool {
        constexpr int NS = decltype(ns_c)::value, OS = decltype(os_c)::value;
        const uint32_t cur = gs->tile_id[it % GS];
        const bool have_new = cur != 0xffffffffu;
        const bool have_old = it >= uint32_t(D - 1) && gs->tile_id[(it - uint32_t(D - 1)) % GS] != 0xffffffffu;
        if (!have_new && !have_old) return false;
        if (have_new) {
            const uint32_t slot = it % D, gslot = it % GS;
            const uint32_t tile_len = gs->len[gslot];
            const uint32_t wall0 = gs->wall[gslot][0], wall1 = gs->wall[gslot][1];  // chunk-last elements in this tile, if any
            const bool full = tile_len == uint32_t(C::TILE);
            FZ_PROF(pf_t = clock64();)
            if (!mbar_wait<2000>(wbar, parity)) *a.scratch.overflow = 3u;
            parity ^= 1u;
            FZ_PROF({ const long long t1 = clock64(); pf_copy += t1 - pf_t; pf_t = t1; ++pf_tiles; })
            // ---- count: lookups (retained), run parity under carry_in = 0, the slice's carry function ----
            bool t_id = true;
            uint32_t t_const = 0, delta = 0, cnt_lane = 0;  // (the lanes' counts are added up once per slice, not per round)
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const uint32_t round_off = slice_off + uint32_t(k * 512);
                const uint32_t off = round_off + uint32_t(lane * 16);  // offset of the lane's segment in the tile
                const uint4 w = *reinterpret_cast<const uint4 *>(slice + k * 512 + lane * 16);
                uint32_t next = __shfl_down_sync(FULL, w.x & 0xffu, 1);
                if (lane == 31) next = slice[k * 512 + 512];
                fz_lookup<0>(tbl_s, w, next, hv[NS][k]);
                fz_lookup<1>(tbl_s, w, next, ov[NS][k]);
                uint32_t valid = 0xFFFFu;
                if (!full) valid = (off + 16 <= tile_len) ? 0xFFFFu : (off < tile_len ? ((1u << (tile_len - off)) - 1u) : 0u);
                if (wall0 - round_off < 512u || wall1 - round_off < 512u) {  // warp-uniform: a wall is in this round
                    // a wall suppresses the pair that starts at the chunk's last element: the raw token goes out there
                    const uint32_t dj0 = wall0 - off, dj1 = wall1 - off;  // >= 16 (or wrapped) in every segment but one
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (dj0 == uint32_t(j) || dj1 == uint32_t(j)) {
                            const uint32_t be = PairsFE::raw_be(w, j);
                            uint32_t &dst = (j & 1) ? ov[NS][k][j >> 2] : hv[NS][k][j >> 2];
                            dst = ((j >> 1) & 1) ? ((dst & 0x0000ffffu) | (be << 16)) : ((dst & 0xffff0000u) | be);
                        }
                    }
                }
                const uint32_t m = fz_membership(hv[NS][k], ov[NS][k]) & valid;
                const uint32_t lead = __clz(~(m << 16));  // ones at the top of the segment
                const uint32_t nid = ~__ballot_sync(FULL, m == 0xFFFFu);
                const uint32_t cob = __ballot_sync(FULL, (lead & 1u) != 0);
                const uint32_t c_round0 = t_id ? 0u : t_const;
                const uint32_t l_nid = nid & lt_mask;
                const uint32_t cin0 = l_nid ? ((cob >> (31 - __clz(l_nid))) & 1u) : c_round0;
                const uint32_t st = start_bits(m, cin0);
                const uint32_t e = valid & ~((st << 1) | cin0);
                const uint32_t cnt = __popc(e);
                em[NS][k] = e;
                cnt_lane += cnt;
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
            const uint32_t cnt0 = __reduce_add_sync(FULL, cnt_lane);
            // every lane has read its share of the buffer: the slice of the next tile may land in it
            const uint32_t nxt_tile = gs->tile_id[(it + 1) % GS];
            if (nxt_tile != 0xffffffffu) fz_warp_copy<C>(a, nxt_tile, wg, slice, wbar, lane);
            else __syncwarp();
            uint32_t last = 0;
            if (lane == 0) {
                gs->fn_flags[slot][wg] = (t_id ? 1u : 0u) | (t_const << 1) | (delta << 2);
                gs->fn_cnt[slot][wg] = cnt0;
                __threadfence_block();
                const uint32_t order = atomicAdd(&gs->arrived[slot], 1u);
                last = (order == uint32_t(WG - 1)) ? 1u : 0u;
                FZ_PROF(if (order == 0u) gs->t_first[slot] = clock64();)
            }
            last = __shfl_sync(FULL, last, 0);
            if (last) {
                // the last worker to finish composes the WG functions and publishes the tile's function right away (no
                // detour through the chain warp: other CTAs' look-backs are waiting for it), then hands the tile over
                __threadfence_block();
                if (lane == 0) gs->arrived[slot] = 0u;
                const FzPending p = fz_chain_publish<WG>(gs, desc, cur, gs->flags[gslot], slot, lane);
                if (lane < WG) {
                    gs->ex_flags[slot][lane] = p.ex.id | (p.ex.cst << 1) | (p.ex.delta << 2);
                    gs->ex_cnt[slot][lane] = p.ex.cnt0;
                }
                if (lane == 0) {
                    gs->tf_flags[slot] = p.tf.id | (p.tf.cst << 1) | (p.tf.delta << 2) | (p.starts << 3);
                    gs->tf_cnt[slot] = p.tf.cnt0;
                }
                __syncwarp();
                FZ_PROF(if (lane == 0) gs->t_pub[slot] = clock64();)
                if (lane == 0) mbar_arrive(bar_counted + 8 * slot);
            }
            FZ_PROF({ const long long t1 = clock64(); pf_count += t1 - pf_t; pf_t = t1; })
        }

        // ---- emit (D-1 tiles behind): compaction of the retained tokens of the whole slice, streamed out in words ----
        if (have_old) {
            const uint32_t jt = it - uint32_t(D - 1);
            const uint32_t slot = jt % D, gslot = jt % GS;
            const uint32_t old_len = gs->len[gslot];
            const uint32_t old_wall0 = gs->wall[gslot][0], old_wall1 = gs->wall[gslot][1];
            const bool old_full = old_len == uint32_t(C::TILE);
            FZ_PROF(pf_t = clock64();)
            if (!mbar_wait<2000>(bar_resolved + 8 * slot, (jt / D) & 1u)) *a.scratch.overflow = 3u;
            FZ_PROF({ const long long t1 = clock64(); pf_chain += t1 - pf_t; pf_t = t1; })
            const unsigned long long rv = gs->res[slot][wg];
            const uint32_t slice_carry = uint32_t(rv >> 63);
            const unsigned long long rel0 = rv & ~R_CARRY;  // tokens of the launch in front of the slice
            const unsigned long long abs0 = rel0 + a.out_base_tokens;
            // logical token 0 of the staging line corresponds to a.out[wpos], a 16-byte boundary of the output; the first
            // `head` slots belong to the slice in front of this one (they hold nothing and are not written here)
            const unsigned long long wpos = abs0 & ~7ull;
            const uint32_t head = uint32_t(abs0 & 7ull);
            if (slice_carry != 0u) {
                // the lanes in front of the slice's first non-identity segment see carry_in = 1 (the count assumed 0)
                bool dep = true;
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    if (dep) {  // warp-uniform
                        const uint32_t off = slice_off + uint32_t(k * 512 + lane * 16);
                        uint32_t valid = 0xFFFFu;
                        if (!old_full) valid = (off + 16 <= old_len) ? 0xFFFFu : (off < old_len ? ((1u << (old_len - off)) - 1u) : 0u);
                        const uint32_t m = fz_membership(hv[OS][k], ov[OS][k]) & valid;
                        const uint32_t nid = ~__ballot_sync(FULL, m == 0xFFFFu);
                        if ((nid & lt_mask) == 0u) {
                            const uint32_t st1 = start_bits(m, 1u);
                            em[OS][k] = valid & ~((st1 << 1) | 1u);
                        }
                        if (nid) dep = false;
                    }
                }
            }
            // token counts of all rounds at once: two rounds per scanned word
            uint32_t pk[R / 2];
#pragma unroll
            for (int h = 0; h < R / 2; ++h) pk[h] = __popc(em[OS][2 * h]) | (__popc(em[OS][2 * h + 1]) << 16);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                for (int h = 0; h < R / 2; ++h) {
                    const uint32_t tq = __shfl_up_sync(FULL, pk[h], d);
                    if (lane >= d) pk[h] += tq;
                }
            }
            uint32_t pos[R];  // tokens of the slice in front of the lane's segment of round k
            uint32_t have = 0;  // tokens of the slice
#pragma unroll
            for (int h = 0; h < R / 2; ++h) {
                const uint32_t tot = __shfl_sync(FULL, pk[h], 31);
                pos[2 * h] = have + (pk[h] & 0xffffu) - __popc(em[OS][2 * h]);
                have += tot & 0xffffu;
                pos[2 * h + 1] = have + (pk[h] >> 16) - __popc(em[OS][2 * h + 1]);
                have += tot >> 16;
            }
            if (a.chunk_ends != nullptr && (old_wall0 & old_wall1) != FZ_NO_WALL) {  // chunks that end in this tile
                const unsigned long long ck0 = gs->wall_ck[gslot][0], ck1 = gs->wall_ck[gslot][1];
#pragma unroll
                for (int k = 0; k < R; ++k) {  // their output ends behind the wall's token
                    const uint32_t off = slice_off + uint32_t(k * 512 + lane * 16);
                    const uint32_t dj0 = old_wall0 - off, dj1 = old_wall1 - off;  // 0 .. 15 in the wall's lane only
                    if (dj0 < 16u) a.chunk_ends[ck0] = a.chunk_ends_base + 2ull * (rel0 + pos[k] + __popc(em[OS][k] & ((2u << dj0) - 1u)));
                    if (dj1 < 16u) a.chunk_ends[ck1] = a.chunk_ends_base + 2ull * (rel0 + pos[k] + __popc(em[OS][k] & ((2u << dj1) - 1u)));
                }
            }
            const bool fits = (abs0 + have <= a.out_cap_tokens);
            if (!fits && lane == 0) *a.scratch.overflow = 1u;
            // dense slice: every lane emits exactly the 8 tokens of one parity in every round
            uint32_t x_or = 0, x_and = 0xFFFFu;
#pragma unroll
            for (int k = 0; k < R; ++k) { x_or |= em[OS][k] ^ 0x5555u; x_and &= em[OS][k] ^ 0x5555u; }
            const bool dense0 = __all_sync(FULL, x_or == 0u), dense1 = __all_sync(FULL, x_and == 0xFFFFu);
            if (fits && (dense0 || dense1) && head == 0u) {
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const uint32_t *tv = dense0 ? hv[OS][k] : ov[OS][k];
                    stg_stream_v4(a.out + abs0 + size_t(k) * 256 + size_t(lane) * 8, make_uint4(tv[0], tv[1], tv[2], tv[3]));
                }
            } else if (fits) {
                // R independent chains of 2-byte stores (position-major so that they interleave), store and cursor bump
                // under the predicate "this position emits" (default).  -DBLT_FZ_UNPRED_STS: the stores are NOT
                // predicated: the token of a position that emits nothing lands in the slot of the lane's next emitting
                // position (no two neighbours are both silent) and is overwritten by it; only position 15, whose
                // successor belongs to the next lane, is predicated; silent positions behind the end of the input pile
                // up in the scratch slot behind the last token (measured slower: more lanes per store, more conflicts).
#ifndef BLT_FZ_NO_BULK_FLUSH
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous tile's bulk copy has read the line
                __syncwarp();
#endif
                uint32_t sp[R];
#pragma unroll
                for (int k = 0; k < R; ++k) sp[k] = stage_s + 2u * (head + pos[k]);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
#pragma unroll
                    for (int k = 0; k < R; ++k) {
                        const uint32_t v = (j & 1) ? ov[OS][k][j >> 2] : hv[OS][k][j >> 2];
                        const uint32_t tok = ((j >> 1) & 1) ? shr_fma(v, 1u << 16) : v;
                        const uint32_t bit = em[OS][k] & (1u << j);
                        uint32_t phys = sp[k];
                        if (kStageSwizzle)
                            asm("{\n\t.reg .b32 t;\n\t"
                                "mul.hi.u32 t, %1, 0x20000000;\n\t"
                                "and.b32 t, t, 0x70;\n\t"
                                "xor.b32 %0, t, %1;\n\t}"
                                : "=r"(phys)
                                : "r"(sp[k]));
#ifndef BLT_FZ_UNPRED_STS
                        if (j < 15 && !kStageSwizzle) {
                            // store and cursor bump under one predicate made by the mask test itself (the same three
                            // instructions as the unpredicated chain; fewer lanes per store, fewer bank conflicts)
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.u32 p, %2, 0;\n\t"
                                "@p st.shared.u16 [%0], %1;\n\t"
                                "@p add.u32 %0, %0, 2;\n\t}"
                                : "+r"(sp[k])
                                : "h"(uint16_t(tok)), "r"(bit)
                                : "memory");
                        } else
#endif
                        if (j < 15) {
                            asm volatile("st.shared.u16 [%0], %1;" ::"r"(phys), "h"(uint16_t(tok)) : "memory");
                            if (j == 0) asm("mad.lo.u32 %0, %1, 2, %0;" : "+r"(sp[k]) : "r"(bit));
                            else if (j == 1) sp[k] += bit;
                            else asm("mad.hi.u32 %0, %1, %2, %0;" : "+r"(sp[k]) : "r"(bit), "r"(1u << (33 - j)));
                        } else {
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.u32 p, %2, 0;\n\t"
                                "@p st.shared.u16 [%0], %1;\n\t}" ::"r"(phys), "h"(uint16_t(tok)), "r"(bit)
                                : "memory");
                        }
                    }
                }
                __syncwarp();
                // tokens head .. head+have-1 of the line go to a.out[wpos + head ...]: whole 16-byte vectors, and the
                // tokens of the two partial vectors at the ends one by one (lanes 0-7: first vector, lanes 8-15: last)
                const uint32_t end = head + have;
                const uint32_t v_first = head != 0u ? 1u : 0u, v_end = end >> 3;
                unsigned char *gout = reinterpret_cast<unsigned char *>(a.out + wpos);
                {
                    const uint32_t L = lane < 8 ? uint32_t(lane) : (8u * v_end + uint32_t(lane) - 8u);
                    const bool mine = lane < 8 ? (head != 0u && L >= head && L < end) : (lane < 16 && L >= head && L < end);
                    if (mine) {
                        uint32_t tv;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(tv) : "r"(stage_swz(stage_s + 2u * L)) : "memory");
                        *reinterpret_cast<uint16_t *>(gout + 2u * L) = uint16_t(tv);
                    }
                }
#ifndef BLT_FZ_NO_BULK_FLUSH
                // the whole vectors leave through the bulk-copy unit (shared -> global, one instruction by one lane): no
                // LDS.128 / STG.128 loop, no wavefronts of the shared-memory load pipe.  The lanes' generic-proxy stores
                // are fenced towards the async proxy first; the line is not written again before the copy has read it
                // (wait_group.read at the top of the next compaction).
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0 && v_end > v_first) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gout + 16u * v_first),
                                 "r"(stage_s + 16u * v_first), "r"(16u * (v_end - v_first))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
#else
#pragma unroll 2
                for (uint32_t v0 = 0; v0 < v_end; v0 += 32u) {  // warp-uniform trip count
                    const uint32_t v = v0 + uint32_t(lane);
                    if (v >= v_first && v < v_end) {
                        uint4 q;
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                                     : "r"(stage_swz(stage_s + 16u * v))
                                     : "memory");
                        stg_stream_v4(gout + 16u * v, q);
                    }
                }
                __syncwarp();
#endif
            }
            FZ_PROF({ const long long t1 = clock64(); pf_emit += t1 - pf_t; pf_t = t1; })
        }
        return true;
    };

    if constexpr (D == 2) {
        // ONE copy of the tile body - count into set 1, emit from set 0, then move set 1 to set 0 (36 register moves per
        // tile) - instead of D copies with the roles of the register sets swapped: 5 064 instead of 8 024 instructions,
        // measured 1 % faster (instruction caches)
        for (uint32_t it = 0;; ++it) {
            if (!iteration(FzIC<1>{}, FzIC<0>{}, it)) break;
#pragma unroll
            for (int k = 0; k < R; ++k) {
#pragma unroll
                for (int q = 0; q < 4; ++q) { hv[0][k][q] = hv[1][k][q]; ov[0][k][q] = ov[1][k][q]; }
                em[0][k] = em[1][k];
            }
        }
    } else {
        for (uint32_t it = 0;; it += uint32_t(D)) {
            if (!iteration(FzIC<0>{}, FzIC<1 % D>{}, it)) break;
            if (!iteration(FzIC<1>{}, FzIC<2 % D>{}, it + 1)) break;
            if constexpr (D == 4) {
                if (!iteration(FzIC<2>{}, FzIC<3>{}, it + 2)) break;
                if (!iteration(FzIC<3>{}, FzIC<0>{}, it + 3)) break;
            }
        }
    }
#ifndef BLT_FZ_NO_BULK_FLUSH
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the CTA's shared memory outlives its bulk copies
    __syncwarp();
#endif
    FZ_PROF(if (lane == 0 && blockIdx.x < 256) {
        unsigned long long *pp = g_fz_prof + (size_t(blockIdx.x) * 32 + wg) * 8;
        uint32_t smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        pp[0] = pf_copy; pp[1] = pf_count; pp[2] = pf_chain; pp[3] = pf_emit; pp[4] = pf_tiles; pp[5] = smid;
    })
}

// One per device: the event behind the most recent fused launch (see FusedLaunch::launch).
struct FusedChain {
    std::mutex mu;
    cudaEvent_t ev = nullptr;
};
inline FusedChain &fused_chain(int dev) {
    static FusedChain chains[kMaxDevices];
    return chains[dev];
}

template <int WG, int R, int D>
struct FusedLaunch {
    using C = FusedCfg<WG, R, D>;
    static cudaError_t configure(int dev) {
        static std::atomic<bool> configured[kMaxDevices];
        if (dev >= kMaxDevices) return cudaErrorInvalidDevice;
        if (!configured[dev].load(std::memory_order_acquire)) {
            cudaError_t err = cudaFuncSetAttribute(fused_sweep_kernel<WG, R, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
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
        // The look-back makes CTAs wait for tiles of OTHER CTAs of the same launch, which is only safe while all of a
        // launch's CTAs are resident (grid <= SM count, one CTA per SM).  Two fused launches on different streams could
        // each get a part of the SMs and wait for their own unscheduled CTAs forever, so fused launches of one device
        // are chained: each waits (on the device, through an event) for the one enqueued before it, whatever its stream.
        // Any other kernel may run beside a fused launch: it ends without waiting for anybody and frees its SMs.
        FusedChain &fc = fused_chain(dev);
        std::lock_guard<std::mutex> lk(fc.mu);
        static const bool no_chain = getenv("BLT_FZ_NO_CHAIN") != nullptr;  // test knob: shows the hazard the chain removes
        if (fc.ev == nullptr) {
            err = cudaEventCreateWithFlags(&fc.ev, cudaEventDisableTiming);
            if (err != cudaSuccess) return err;
        } else if (!no_chain) {
            err = cudaStreamWaitEvent(stream, fc.ev, 0);
            if (err != cudaSuccess) return err;
        }
        fused_sweep_kernel<WG, R, D><<<dim3(unsigned(grid)), dim3(C::THREADS), C::SMEM, stream>>>(
            a, d_table, reinterpret_cast<unsigned long long *>(a.scratch.meta), uint32_t(tiles), (unsigned long long)chunk);
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
        return cudaEventRecord(fc.ev, stream);
    }
};
