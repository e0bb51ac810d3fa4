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
    for (; wk.cur < wk.end; ++wk.cur) {
        uint32_t nv;
        const uint4 w = detok_load(a, wk.cur, lane, &nv);
        bytes += nv + __popc(detok_wide_mask(w));  // absent tokens are zero: never wide
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
    for (; wk.cur < wk.end; ++wk.cur) {
        uint32_t nv;
        const uint4 w = detok_load(a, wk.cur, lane, &nv);
        const uint32_t wide = detok_wide_mask(w);
        const uint32_t cnt = nv + __popc(wide);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        unsigned char *sp = stage + pend + (incl - cnt);
        const uint32_t words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (uint32_t(k) < nv) {
                const uint32_t half = (words[k >> 1] >> (16 * (k & 1))) & 0xffffu;   // stored: hi | lo << 8
                const uint32_t id = __byte_perm(half, 0, 0x4401);                    // the token
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
    // the tail of the range: bytes [head, pend) of a vector shared with the next warp's range
    for (uint32_t k = head + lane; k < pend; k += 32) a.out[wpos + k] = stage[k];
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
