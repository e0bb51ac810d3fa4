// pairhist.cuh -- adjacent-byte-pair histogram of a corpus (merges "training").
//
// The reference only consumes merges files (config_loader.rs:14-46); the tables of its benchmark configs are
// "the k most frequent adjacent byte pairs" of a sample, which needs exactly this histogram:
// counts[b0 << 8 | b1] = number of i with in[i] = b0, in[i+1] = b1.
// One persistent CTA per SM keeps all 65 536 counters in shared memory as u16 (128 KiB), bumped with 32-bit
// shared-memory atomics on the containing word.  The atomic returns the old word: a lane that sees its counter at
// 16 384 or more raises a flag, the CTA looks at the flag every two rounds (32 768 positions, so a counter that was
// below 16 384 at the last look cannot wrap before the next one) and only then adds its non-zero counters to the
// global u64 histogram and clears them: uniform data never flushes before the end (a flush there is 65 536 global
// atomics), text flushes when its hottest pair gets there.  Every position is read once.
// Included by kernels.cu inside its anonymous namespace.
#pragma once

constexpr int kHistRoundsPerLook = 2;      // 2 rounds x 1024 threads x 16 positions = 32 768
constexpr uint32_t kHistFlushAt = 16384u;  // 16 383 + 32 768 + 16 384 < 65 536

__global__ void __launch_bounds__(kCtaThreads, 1)
pair_hist_kernel(const unsigned char *__restrict__ in, unsigned long long n, unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem);  // 32 768 words = 65 536 u16 counters
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) reinterpret_cast<uint4 *>(hist)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned long long n_pairs = n ? n - 1 : 0;
    const unsigned long long cta_span = (unsigned long long)kCtaThreads * 16;     // positions per CTA-round
    unsigned long long round = blockIdx.x;
    const unsigned long long n_rounds = (n_pairs + cta_span - 1) / cta_span;
    int since_look = 0;
    bool dirty = false;
    uint32_t hot = 0;  // this lane saw a counter at kHistFlushAt or more
    auto flush = [&]() {
        __syncthreads();
        for (int i = threadIdx.x; i < 8192; i += blockDim.x) {
            const uint4 v = reinterpret_cast<uint4 *>(hist)[i];
            if ((v.x | v.y | v.z | v.w) == 0u) continue;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t key = uint32_t(i) * 8u + uint32_t(k) * 2u;
                if (w[k] & 0xffffu) atomicAdd(counts + key, (unsigned long long)(w[k] & 0xffffu));
                if (w[k] >> 16) atomicAdd(counts + key + 1, (unsigned long long)(w[k] >> 16));
            }
            reinterpret_cast<uint4 *>(hist)[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
    };
    for (; round < n_rounds; round += gridDim.x) {
        const unsigned long long p0 = round * cta_span + (unsigned long long)threadIdx.x * 16;  // first position of the lane
        uint4 w = make_uint4(0, 0, 0, 0);
        uint32_t valid = 0;  // number of pair positions of this lane that exist (0..16)
        if (p0 + 16 <= n) {
            w = ldg_stream_v4(in + p0);
            valid = (p0 + 16 <= n_pairs) ? 16u : uint32_t(n_pairs - p0);
        } else if (p0 < n) {
            uint32_t tmp[4] = {0, 0, 0, 0};
            for (unsigned j = 0; p0 + j < n; ++j) tmp[j >> 2] |= uint32_t(in[p0 + j]) << (8 * (j & 3));
            w = make_uint4(tmp[0], tmp[1], tmp[2], tmp[3]);
            valid = p0 < n_pairs ? uint32_t(n_pairs - p0) : 0u;
        }
        uint32_t next = __shfl_down_sync(FULL, w.x & 0xffu, 1);
        if (lane == 31) next = (valid == 16u) ? uint32_t(in[p0 + 16]) : 0u;
        const uint32_t words[5] = {w.x, w.y, w.z, w.w, next};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            // bytes j and j+1 of the lane's window
            const uint32_t lo = __funnelshift_r(words[j >> 2], words[(j >> 2) + 1], 8 * (j & 3));
            const uint32_t key = ((lo & 0xffu) << 8) | ((lo >> 8) & 0xffu);          // b0 << 8 | b1
            if (uint32_t(j) < valid) {
                const uint32_t old = atomicAdd(hist + (key >> 1), 1u << (16 * (key & 1u)));
                hot |= ((old >> (16 * (key & 1u))) & 0xffffu) >= kHistFlushAt ? 1u : 0u;
            }
        }
        dirty = true;
        if (++since_look == kHistRoundsPerLook) {
            since_look = 0;
            if (__syncthreads_or(int(hot))) { flush(); hot = 0; dirty = false; }
        }
    }
    if (dirty) flush();
}

cudaError_t launch_pair_hist_impl(const unsigned char *d_in, size_t n, unsigned long long *d_counts, bool zero_first,
                                  cudaStream_t stream) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev >= kMaxDevices) return cudaErrorInvalidDevice;
    static std::atomic<bool> configured[kMaxDevices];
    if (!configured[dev].load(std::memory_order_acquire)) {
        err = cudaFuncSetAttribute(pair_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
        if (err != cudaSuccess) return err;
        configured[dev].store(true, std::memory_order_release);
    }
    if (zero_first) {
        err = cudaMemsetAsync(d_counts, 0, 65536 * sizeof(unsigned long long), stream);
        if (err != cudaSuccess) return err;
    }
    if (n < 2) return cudaSuccess;
    const size_t n_rounds = (n - 1 + size_t(kCtaThreads) * 16 - 1) / (size_t(kCtaThreads) * 16);
    size_t grid = std::min<size_t>(n_rounds, size_t(sm_count(dev)));
    pair_hist_kernel<<<dim3(unsigned(grid)), dim3(kCtaThreads), 131072, stream>>>(d_in, n, d_counts);
    return cudaGetLastError();
}
