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

// bit k <-> token k of the lane is >= 256 (its high byte, the low byte of the stored half-word, is non-zero)
__device__ __forceinline__ uint32_t detok_wide_mask(const uint4 &w) {
    return PairsFE::nonzero_bytes4(__byte_perm(w.x, w.y, 0x6420)) | (PairsFE::nonzero_bytes4(__byte_perm(w.z, w.w, 0x6420)) << 4);
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
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            nvq[r] = 0;
            wq[r] = make_uint4(0, 0, 0, 0);
            if (wk.cur + r < wk.end) wq[r] = detok_load(a, wk.cur + r, lane, &nvq[r]);
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) bytes += nvq[r] + __popc(detok_wide_mask(wq[r]));  // absent tokens are zero: never wide
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

template <bool HOLES>
__global__ void __launch_bounds__(kCtaThreads, 1) detok_emit_kernel(const DetokArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    {   // id -> l | r << 8, and (HOLES) the bitmap of ids that exist
        const uint4 *src = reinterpret_cast<const uint4 *>(a.table);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        const int n16 = (kPairTableEntries * 2 + (HOLES ? 8192 : 0)) / 16;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    if (*reinterpret_cast<const volatile uint32_t *>(a.scratch.overflow) != 0u) return;  // the scan found it does not fit
    const uint16_t *dec = reinterpret_cast<const uint16_t *>(smem);
    const uint32_t *exists = reinterpret_cast<const uint32_t *>(smem + kPairTableEntries * 2);
    const int lane = threadIdx.x & 31;
    unsigned char *stage = smem + kPairTableEntries * 2 + 8192 + size_t(threadIdx.x >> 5) * kDetokStageBytes;
    const long long warp = (long long)blockIdx.x * (kCtaThreads / 32) + (threadIdx.x >> 5);
    DetokWalk wk;
    wk.init(a.n_tok, warp, (long long)gridDim.x * (kCtaThreads / 32));
    if (wk.cur >= wk.end || warp >= kMaxRanges) return;
    const unsigned long long base = a.scratch.tile_status[warp];
    // stage[0 .. pend) holds bytes not yet written; stage[0] is a.out[wpos], wpos a multiple of 16.  The first
    // `head` bytes of the very first vector belong to the previous warp's range.
    unsigned long long wpos = base & ~15ull;
    uint32_t pend = uint32_t(base & 15ull);
    uint32_t head = pend;
    bool bad = false;
    uint32_t idmax = 0;    // packed running maximum of the ids seen by the all-merged fast path (checked at the end)
    constexpr int RB = 4;  // rounds loaded together: four independent 16-byte loads in flight per lane
    for (; wk.cur < wk.end; wk.cur += RB) {
      uint4 wq[RB];
      uint32_t nvq[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
          nvq[r] = 0;
          wq[r] = make_uint4(0, 0, 0, 0);
          if (wk.cur + r < wk.end) wq[r] = detok_load(a, wk.cur + r, lane, &nvq[r]);
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (wk.cur + r >= wk.end) break;
        const uint4 w = wq[r];
        const uint32_t nv = nvq[r];
        const uint32_t wide = detok_wide_mask(w);
        const uint32_t cnt = nv + __popc(wide);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        const uint32_t words[4] = {w.x, w.y, w.z, w.w};
        bool direct = false;
        if (pend == 0 && total == 512u) {
            // ---- every token of the round is a merged id and the output is vector-aligned: 8 lookups, one
            // 16-byte store per lane straight from registers (the steady state on a full table's output) ----
            uint32_t o4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t ids = __byte_perm(words[k], 0, 0x2301);               // both halves byte-swapped
                idmax = __vmaxu2(idmax, ids);
                const uint32_t e0 = *reinterpret_cast<const uint16_t *>(smem + ((ids << 1) & 0x1FFFEu));
                const uint32_t e1 = *reinterpret_cast<const uint16_t *>(smem + ((ids >> 15) & 0x1FFFEu));
                if (HOLES) {
                    const uint32_t i0 = ids & 0xffffu, i1 = ids >> 16;
                    bad = bad || (((exists[i0 >> 5] >> (i0 & 31)) & (exists[i1 >> 5] >> (i1 & 31)) & 1u) == 0u);
                }
                o4[k] = e0 | (e1 << 16);
            }
            stg_stream_v4(a.out + wpos + 16ull * lane, make_uint4(o4[0], o4[1], o4[2], o4[3]));
            wpos += 512u;
            continue;
        }
        if (__all_sync(FULL, nv == 8)) {
            // ---- full round: the lane's bytes are packed in registers (a tree of shifts by the widths) ----
            // (branch-free: the table is read for narrow tokens too and the result dropped)
            uint32_t v[8], wd[8], badm = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t half = (words[k >> 1] >> (16 * (k & 1))) & 0xffffu;   // stored: hi | lo << 8
                const uint32_t id = __byte_perm(half, 0, 0x4401);                    // the token
                const uint32_t is_wide = (wide >> k) & 1u;
                const uint32_t e = dec[id];
                wd[k] = 8u + 8u * is_wide;
                v[k] = is_wide ? e : (half >> 8);
                uint32_t unknown = (id >= a.limit) ? 1u : 0u;
                if (HOLES) unknown |= ~(exists[id >> 5] >> (id & 31)) & 1u;
                badm |= unknown & is_wide;
            }
            bad = bad || badm != 0u;
            const uint32_t c01 = v[0] | (v[1] << wd[0]), c23 = v[2] | (v[3] << wd[2]);
            const uint32_t c45 = v[4] | (v[5] << wd[4]), c67 = v[6] | (v[7] << wd[6]);
            const uint32_t w01 = wd[0] + wd[1], w45 = wd[4] + wd[5];                 // 16..32 bits
            // d0 = c01 | c23 << w01, d1 = c45 | c67 << w45 (64-bit each, as two words; clamped funnel shifts)
            const uint32_t d0l = c01 | __funnelshift_lc(0u, c23, w01), d0h = __funnelshift_lc(c23, 0u, w01);
            const uint32_t d1l = c45 | __funnelshift_lc(0u, c67, w45), d1h = __funnelshift_lc(c67, 0u, w45);
            const uint32_t sh2 = w01 + wd[2] + wd[3] - 32u;                          // S = d0 | d1 << (32 + sh2), sh2 in 0..32
            const uint32_t s0 = d0l, s1 = d0h | __funnelshift_lc(0u, d1l, sh2);
            const uint32_t s2 = __funnelshift_lc(d1l, d1h, sh2), s3 = __funnelshift_lc(d1h, 0u, sh2);
            if (pend == 0 && total == 256u) {
                // every token of the round is a plain byte and the output is vector-aligned: straight out
                direct = true;
                *reinterpret_cast<uint2 *>(a.out + wpos + 8ull * lane) = make_uint2(s0, s1);
                wpos += total;
            } else {
                // place the lane's cnt bytes at byte offset o of the staging line with 4-byte stores: shift by
                // o & 3, take the previous lane's incomplete last word into my first one, store my complete words
                const uint32_t o = pend + (incl - cnt);
                const uint32_t sh = (o & 3u) * 8u;
                uint32_t t0 = s0 << sh;
                const uint32_t t1 = __funnelshift_l(s0, s1, sh), t2 = __funnelshift_l(s1, s2, sh);
                const uint32_t t3 = __funnelshift_l(s2, s3, sh), t4 = __funnelshift_l(s3, 0u, sh);
                const uint32_t cw = ((o + cnt) >> 2) - (o >> 2);                     // complete words: 2..4
                const bool ragged_end = ((o + cnt) & 3u) != 0u;
                const uint32_t my_tail = !ragged_end ? 0u : (cw == 2 ? t2 : cw == 3 ? t3 : t4);
                uint32_t prev = __shfl_up_sync(FULL, my_tail, 1);
                uint32_t *sw = reinterpret_cast<uint32_t *>(stage) + (o >> 2);
                if (lane == 0) prev = sh ? (sw[0] & ((1u << sh) - 1u)) : 0u;          // left by the previous round
                t0 |= prev;
                sw[0] = t0;
                sw[1] = t1;
                if (cw > 2) sw[2] = t2;
                if (cw > 3) sw[3] = t3;
                if (lane == 31 && ragged_end) sw[cw] = my_tail;
            }
        } else {
            // ---- the ragged last round of the stream: byte stores ----
            unsigned char *sp = stage + pend + (incl - cnt);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (uint32_t(k) < nv) {
                    const uint32_t half = (words[k >> 1] >> (16 * (k & 1))) & 0xffffu;
                    const uint32_t id = __byte_perm(half, 0, 0x4401);
                    if ((wide >> k) & 1u) {
                        const uint32_t e = dec[id];
                        bad = bad || id >= a.limit || (HOLES && ((exists[id >> 5] >> (id & 31)) & 1u) == 0u);
                        sp[0] = static_cast<unsigned char>(e);
                        sp[1] = static_cast<unsigned char>(e >> 8);
                        sp += 2;
                    } else {
                        sp[0] = static_cast<unsigned char>(id);
                        sp += 1;
                    }
                }
            }
        }
        if (direct) continue;
        __syncwarp();
        // flush the whole 16-byte vectors, keep the leftover (< 16 bytes) at the front of the line
        const uint32_t have = pend + total;
        const uint32_t nvec = have >> 4;
        for (uint32_t v = lane; v < nvec; v += 32) {
            if (v == 0 && head != 0) {
                for (uint32_t k = head; k < 16; ++k) a.out[wpos + k] = stage[k];
            } else {
                stg_stream_v4(a.out + wpos + 16ull * v, *reinterpret_cast<const uint4 *>(stage + 16 * v));
            }
        }
        const uint32_t rem = have & 15u;
        unsigned char keep = 0;
        if (nvec != 0 && uint32_t(lane) < rem) keep = stage[16 * nvec + lane];
        __syncwarp();
        if (nvec != 0) {
            if (uint32_t(lane) < rem) stage[lane] = keep;
            head = 0;
            wpos += 16ull * nvec;
            pend = rem;
        } else {
            pend = have;
        }
        __syncwarp();
      }
    }
    // the tail of the range: bytes [head, pend) of a vector shared with the next warp's range
    for (uint32_t k = head + lane; k < pend; k += 32) a.out[wpos + k] = stage[k];
    bad = bad || (idmax & 0xffffu) >= a.limit || (idmax >> 16) >= a.limit;
    if (__any_sync(FULL, bad) && lane == 0) reinterpret_cast<uint32_t *>(a.scratch.ctrl)[6] = 1u;  // "unknown token"
}

cudaError_t launch_detok_impl(const DetokArgs &a, cudaStream_t stream) {
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
        configured[dev].store(true, std::memory_order_release);
    }
    if (a.scratch.max_tiles < size_t(2 * kMaxRanges)) return cudaErrorInvalidValue;
    err = cudaMemsetAsync(a.scratch.ctrl, 0, 256, stream);
    if (err != cudaSuccess) return err;
    const size_t n_rounds = (a.n_tok + kDetokRoundTokens - 1) / kDetokRoundTokens;
    const size_t warps_per_cta = kCtaThreads / 32;
    size_t grid = (n_rounds + warps_per_cta - 1) / warps_per_cta;
    if (grid > size_t(sm_count(dev))) grid = size_t(sm_count(dev));
    if (grid > size_t(kMaxRanges) / warps_per_cta) grid = size_t(kMaxRanges) / warps_per_cta;
    if (grid == 0) grid = 1;
    detok_count_kernel<<<dim3(unsigned(grid)), dim3(kCtaThreads), 0, stream>>>(a);
    detok_scan_kernel<<<1, kCtaThreads, 0, stream>>>(a, int(grid * warps_per_cta));
    if (a.holes) detok_emit_kernel<true><<<dim3(unsigned(grid)), dim3(kCtaThreads), kDetokSmem, stream>>>(a);
    else detok_emit_kernel<false><<<dim3(unsigned(grid)), dim3(kCtaThreads), kDetokSmem, stream>>>(a);
    return cudaGetLastError();
}
