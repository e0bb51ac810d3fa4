// pipeline.cu -- host side of the chunk pipeline: the in-memory form (blt_tokenize_host) and the
// file-to-file form (blt_run_tokenizer) of run_mmap_pipeline / run_stream_pipeline
// (blt_core/src/pipeline.rs:56-240).  Chunks are cut at fixed offsets k*C from the start of the input
// (pipeline.rs:73-81), flow through S slots (H2D copy stream -> compute stream -> D2H copy stream)
// and are written strictly in chunk order (pipeline.rs:153-168).  With several GPUs the chunks are dealt
// round-robin (chunk k -> GPU k mod G, blt_file_chunk_device) and each device runs its own pipeline; the
// only cross-GPU datum is each chunk's output length (a host-side prefix), so nothing is exchanged
// between devices.
#include "pipeline.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fcntl.h>
#include <functional>
#include <sys/mman.h>
#include <sys/stat.h>
#include <chrono>
#include <thread>
#include <unistd.h>

namespace bltc {

#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return bltc::fail(e__ == cudaErrorMemoryAllocation ? BLT_ERR_NOMEM : BLT_ERR_CUDA,     \
                              std::string(#expr) + ": " + cudaGetErrorString(e__));                \
    } while (0)

constexpr size_t kSlots = 3;

// ---- Pipe -----------------------------------------------------------------------------------------
int Pipe::ensure(size_t chunk_cap, size_t n_slots, bool want_pinned) {
    if (!s_h2d) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
    }
    if (chunk_cap > cap) {
        for (Slot &sl : slots) {
            if (sl.d_in) cudaFree(sl.d_in);
            if (sl.d_out) cudaFree(sl.d_out);
            sl.d_in = sl.d_out = nullptr;
        }
        cap = chunk_cap;
    }
    if (want_pinned && chunk_cap > pin_cap) {
        for (Slot &sl : slots) {
            if (sl.h_in) cudaFreeHost(sl.h_in);
            if (sl.h_out) cudaFreeHost(sl.h_out);
            sl.h_in = sl.h_out = nullptr;
        }
        pin_cap = chunk_cap;
    }
    if (slots.size() < n_slots) slots.resize(n_slots);
    for (size_t i = 0; i < n_slots; ++i) {
        Slot &sl = slots[i];
        if (!sl.ev_h2d) {
            CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_d2h, cudaEventDisableTiming));
            CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_ctrl), 64, cudaHostAllocDefault));
        }
        if (!sl.d_in) {
            CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&sl.d_in), cap + 64));
            CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&sl.d_out), 2 * cap + 64));
        }
        if (want_pinned && !sl.h_in) {
            CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_in), pin_cap, cudaHostAllocDefault));
            CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_out), 2 * pin_cap, cudaHostAllocDefault));
        }
    }
    return BLT_OK;
}

void Pipe::release() {
    for (Slot &sl : slots) {
        if (sl.d_in) cudaFree(sl.d_in);
        if (sl.d_out) cudaFree(sl.d_out);
        if (sl.h_in) cudaFreeHost(sl.h_in);
        if (sl.h_out) cudaFreeHost(sl.h_out);
        if (sl.h_ctrl) cudaFreeHost(sl.h_ctrl);
        if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
        if (sl.ev_d2h) cudaEventDestroy(sl.ev_d2h);
    }
    slots.clear();
    for (cudaEvent_t e : piece_ev) cudaEventDestroy(e);
    piece_ev.clear();
    if (s_h2d) cudaStreamDestroy(s_h2d);
    if (s_comp) cudaStreamDestroy(s_comp);
    if (s_d2h) cudaStreamDestroy(s_d2h);
    s_h2d = s_comp = s_d2h = nullptr;
    ws.release();
    cap = 0;
    pin_cap = 0;
}

}  // namespace bltc

std::unique_ptr<bltc::Pipe> blt_ctx::acquire() {
    std::lock_guard<std::mutex> lk(mu);
    if (!idle.empty()) {
        auto p = std::move(idle.back());
        idle.pop_back();
        return p;
    }
    auto p = std::unique_ptr<bltc::Pipe>(new bltc::Pipe());
    p->device = device;
    return p;
}

void blt_ctx::give_back(std::unique_ptr<bltc::Pipe> p) {
    std::lock_guard<std::mutex> lk(mu);
    idle.push_back(std::move(p));
}

namespace bltc {

namespace {

// A blocking parallel-for over a few helper threads: the host-side copies of the file pipeline (page
// cache -> pinned staging, pinned staging -> output file) are memory-bound single-threaded otherwise.
// The reference spreads the same work over its tokio workers (`--threads`, utils.rs:79-97).
class HostPool {
  public:
    explicit HostPool(size_t helpers) {
        for (size_t i = 0; i < helpers; ++i) threads_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    size_t width() const { return threads_.size() + 1; }
    // runs fn(0..parts-1); the caller takes part in the work; returns when every part is done
    template <class F>
    void run(size_t parts, F fn) {
        if (parts <= 1 || threads_.empty()) { for (size_t i = 0; i < parts; ++i) fn(i); return; }
        std::function<void(size_t)> f = fn;
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &f; next_ = 0; parts_ = parts; pending_ = parts;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void work() {
        for (;;) {
            size_t i;
            const std::function<void(size_t)> *f;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!fn_ || next_ >= parts_) return;
                i = next_++;
                f = fn_;
            }
            (*f)(i);
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_all();
        }
    }
    void loop() {
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || (fn_ && next_ < parts_); });
                if (stop_) return;
            }
            work();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t next_ = 0, parts_ = 0, pending_ = 0;
    bool stop_ = false;
};


// One process-wide pool for callers that hand in pageable buffers (blt_process_chunk from the reference's
// workers): whoever finds it free copies with its helpers, everybody else copies alone.
void shared_par_memcpy(uint8_t *dst, const uint8_t *src, size_t len) {
    static HostPool pool(std::min<size_t>(8, std::max<size_t>(2, std::thread::hardware_concurrency() / 2)) - 1);
    static std::mutex busy;
    constexpr size_t kPiece = size_t(1) << 20;
    if (len >= 4 * kPiece && busy.try_lock()) {
        const size_t parts = std::min(pool.width(), (len + kPiece - 1) / kPiece);
        const size_t per = ((len + parts - 1) / parts + 4095) & ~size_t(4095);
        pool.run(parts, [&](size_t i) {
            const size_t lo = i * per;
            if (lo < len) std::memcpy(dst + lo, src + lo, std::min(per, len - lo));
        });
        busy.unlock();
        return;
    }
    std::memcpy(dst, src, len);
}

bool is_pinned_host(const void *p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// The slot-pipelined loop shared by the in-memory and the file pipelines.
//   fetch(k, slot)      -> host pointer to chunk k's input bytes (may stage into slot.h_in)
//   deliver(k, slot, n) -> consume chunk k's n output bytes, which the D2H stream is writing to
//                          `dst`; called after the copy has been enqueued; returns where chunk k's
//                          bytes must land (see users)
// Chunk order is preserved by construction: completion is processed for k = 0,1,2,... in order.
struct ChunkSource {
    size_t n = 0, chunk = 0;           // units of `chunk` bytes of an n-byte input
    size_t first = 0, count = 0, stride = 1;  // this pipeline owns units first, first+stride, ... (count of them)
    size_t wall = 0;                   // the reference's chunk size inside a unit (0: unit == chunk)
    bool detok = false;                // units are token bytes to detokenize
    std::vector<std::pair<size_t, size_t>> units;  // optional explicit (offset, length) list instead of uniform units
    size_t id_of(size_t i) const { return first + i * stride; }
    size_t len_of(size_t id) const { return units.empty() ? std::min(chunk, n - id * chunk) : units[id].second; }
    size_t off_of(size_t id) const { return units.empty() ? id * chunk : units[id].first; }
};

template <class Fetch, class Sink>
int run_slots(blt_strategy *s, Pipe &pipe, const ChunkSource &src, Fetch fetch, Sink sink) {
    const size_t S = std::min(kSlots, src.count);
    auto issue = [&](size_t i) -> int {
        Slot &sl = pipe.slots[i % S];
        const size_t id = src.id_of(i);
        sl.in_len = src.len_of(id);
        const bool reuse = i >= S;
        if (reuse) CUDA_TRY(cudaStreamWaitEvent(pipe.s_h2d, sl.ev_done, 0));   // d_in free once its kernel ran
        const uint8_t *h = fetch(id, sl);
        CUDA_TRY(cudaMemcpyAsync(sl.d_in, h, sl.in_len, cudaMemcpyHostToDevice, pipe.s_h2d));
        CUDA_TRY(cudaEventRecord(sl.ev_h2d, pipe.s_h2d));
        CUDA_TRY(cudaStreamWaitEvent(pipe.s_comp, sl.ev_h2d, 0));
        if (reuse) CUDA_TRY(cudaStreamWaitEvent(pipe.s_comp, sl.ev_d2h, 0));  // d_out free once copied out
        int rc = src.detok ? run_detok(s, pipe.ws, sl.d_in, sl.in_len, sl.d_out, 2 * pipe.cap, pipe.s_comp, &sl.res)
                           : run_device(s, pipe.ws, sl.d_in, sl.in_len, src.wall, sl.d_out, 2 * pipe.cap, nullptr, pipe.s_comp, &sl.res);
        if (rc) return rc;
        if (sl.res.kind == DeviceResult::IN_SCRATCH)
            CUDA_TRY(cudaMemcpyAsync(sl.h_ctrl, pipe.ws.scratch.ctrl, 32, cudaMemcpyDeviceToHost, pipe.s_comp));
        CUDA_TRY(cudaEventRecord(sl.ev_done, pipe.s_comp));
        return BLT_OK;
    };
    for (size_t i = 0; i < S; ++i) {
        int rc = issue(i);
        if (rc) return rc;
    }
    for (size_t i = 0; i < src.count; ++i) {
        Slot &sl = pipe.slots[i % S];
        CUDA_TRY(cudaEventSynchronize(sl.ev_done));
        if (sl.res.kind == DeviceResult::IN_SCRATCH) {
            int rc = decode_ctrl(sl.h_ctrl, &sl.res);
            if (rc) return rc;
        }
        int rc = sink(src.id_of(i), sl, sl.res.len);  // enqueues the D2H copy on s_d2h and records ev_d2h
        if (rc) return rc;
        if (i + S < src.count) {
            rc = issue(i + S);
            if (rc) return rc;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(pipe.s_d2h));
    return BLT_OK;
}

// ONE unit through pageable caller buffers (blt_process_chunk as the reference's workers call it: a slice of an
// mmap in, a fresh Vec<u8> out).  The unit is moved in pieces so that the host-side staging copies run beside
// the DMA in both directions:  memcpy(piece i) || H2D(piece i-1)  ->  kernel(s) on the whole unit  ->
// D2H(piece i+1) || memcpy(piece i).  The kernel needs the whole unit (run parity crosses pieces), so it sits
// between the two pipelines; at 3 TB/s it is the smallest term.
int run_one_unit_pieced(blt_strategy *s, Pipe &pipe, const ChunkSource &src, const uint8_t *in, uint8_t *out, size_t out_cap,
                        size_t off, bool stage_in, bool stage_out, size_t *out_len) {
    constexpr size_t kPiece = size_t(4) << 20;
    Slot &sl = pipe.slots[0];
    const size_t id = src.id_of(0);
    const size_t len_in = src.len_of(id);
    const uint8_t *base = in + src.off_of(id);
    sl.in_len = len_in;
    for (size_t o = 0; o < len_in; o += kPiece) {
        const size_t l = std::min(kPiece, len_in - o);
        const uint8_t *h = base + o;
        if (stage_in) {
            shared_par_memcpy(sl.h_in + o, base + o, l);
            h = sl.h_in + o;
        }
        CUDA_TRY(cudaMemcpyAsync(sl.d_in + o, h, l, cudaMemcpyHostToDevice, pipe.s_h2d));
    }
    CUDA_TRY(cudaEventRecord(sl.ev_h2d, pipe.s_h2d));
    CUDA_TRY(cudaStreamWaitEvent(pipe.s_comp, sl.ev_h2d, 0));
    int rc = src.detok ? run_detok(s, pipe.ws, sl.d_in, sl.in_len, sl.d_out, 2 * pipe.cap, pipe.s_comp, &sl.res)
                       : run_device(s, pipe.ws, sl.d_in, sl.in_len, src.wall, sl.d_out, 2 * pipe.cap, nullptr, pipe.s_comp, &sl.res);
    if (rc) return rc;
    if (sl.res.kind == DeviceResult::IN_SCRATCH)
        CUDA_TRY(cudaMemcpyAsync(sl.h_ctrl, pipe.ws.scratch.ctrl, 32, cudaMemcpyDeviceToHost, pipe.s_comp));
    CUDA_TRY(cudaEventRecord(sl.ev_done, pipe.s_comp));
    CUDA_TRY(cudaEventSynchronize(sl.ev_done));
    if (sl.res.kind == DeviceResult::IN_SCRATCH) {
        rc = decode_ctrl(sl.h_ctrl, &sl.res);
        if (rc) return rc;
    }
    const size_t len = sl.res.len;
    if (off + len > out_cap) return fail(BLT_ERR_CAPACITY, "output capacity exceeded");
    const size_t n_pieces = (len + kPiece - 1) / kPiece;
    while (pipe.piece_ev.size() < n_pieces) {
        cudaEvent_t e = nullptr;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        pipe.piece_ev.push_back(e);
    }
    for (size_t i = 0; i < n_pieces; ++i) {
        const size_t o = i * kPiece, l = std::min(kPiece, len - o);
        CUDA_TRY(cudaMemcpyAsync(stage_out ? sl.h_out + o : out + off + o, sl.d_out + o, l, cudaMemcpyDeviceToHost, pipe.s_d2h));
        CUDA_TRY(cudaEventRecord(pipe.piece_ev[i], pipe.s_d2h));
    }
    for (size_t i = 0; i < n_pieces; ++i) {
        CUDA_TRY(cudaEventSynchronize(pipe.piece_ev[i]));
        const size_t o = i * kPiece, l = std::min(kPiece, len - o);
        if (stage_out) shared_par_memcpy(out + off + o, sl.h_out + o, l);
    }
    *out_len = off + len;
    return BLT_OK;
}

// Runs the units of `src` through a leased pipe: host `in` -> device -> host `out` (appended from byte `off`).
// stage_in / stage_out: that side is pageable and goes through the pipe's pinned slots.
int run_host_units(blt_strategy *s, const ChunkSource &src, const uint8_t *in, uint8_t *out, size_t out_cap, size_t off,
                   size_t unit_cap, bool stage_in, bool stage_out, size_t *out_len) {
    auto pipe = s->ctx->acquire();
    int rc = pipe->ensure(unit_cap, std::min(kSlots, src.count), stage_in || stage_out);
    // Measured (tools/chunk_api_bench.py, 16 MiB chunks, 16-core box): pieces 4.4 / 6.6 / 14.3 / 19.2 GB/s at 1 / 2 / 8 / 16
    // callers against 5.3 / 8.5 / 15.1 / 17.9 with the unit as one piece: the helper-pool hand-offs per piece cost more
    // than the overlap wins until the callers outnumber the cores.  Off by default; BLT_PIECED=1 turns it on.
    static const bool pieced = getenv("BLT_PIECED") != nullptr;
    if (rc == BLT_OK && src.count == 1 && (stage_in || stage_out) && pieced) {
        size_t total = off;
        rc = run_one_unit_pieced(s, *pipe, src, in, out, out_cap, off, stage_in, stage_out, &total);
        if (rc != BLT_OK) {
            cudaStreamSynchronize(pipe->s_h2d);
            cudaStreamSynchronize(pipe->s_comp);
            cudaStreamSynchronize(pipe->s_d2h);
        }
        s->ctx->give_back(std::move(pipe));
        if (rc == BLT_OK) *out_len = total;
        return rc;
    }
    // staged output is handed over one unit late, so that its D2H overlaps the next unit's staging
    struct Pending { Slot *sl = nullptr; size_t len = 0, at = 0; } pend;
    auto hand_over = [&](Pending &p) -> int {
        if (!p.sl) return BLT_OK;
        CUDA_TRY(cudaEventSynchronize(p.sl->ev_d2h));
        shared_par_memcpy(out + p.at, p.sl->h_out, p.len);
        p.sl = nullptr;
        return BLT_OK;
    };
    if (rc == BLT_OK) {
        rc = run_slots(
            s, *pipe, src,
            [&](size_t k, Slot &sl) -> const uint8_t * {
                if (!stage_in) return in + src.off_of(k);
                shared_par_memcpy(sl.h_in, in + src.off_of(k), src.len_of(k));
                return sl.h_in;
            },
            [&](size_t, Slot &sl, size_t len) -> int {
                if (off + len > out_cap) return fail(BLT_ERR_CAPACITY, "output capacity exceeded");
                if (stage_out) {
                    int w = hand_over(pend);
                    if (w) return w;
                    CUDA_TRY(cudaMemcpyAsync(sl.h_out, sl.d_out, len, cudaMemcpyDeviceToHost, pipe->s_d2h));
                    pend.sl = &sl; pend.len = len; pend.at = off;
                } else {
                    CUDA_TRY(cudaMemcpyAsync(out + off, sl.d_out, len, cudaMemcpyDeviceToHost, pipe->s_d2h));
                }
                CUDA_TRY(cudaEventRecord(sl.ev_d2h, pipe->s_d2h));
                off += len;
                return BLT_OK;
            });
        if (rc == BLT_OK) rc = hand_over(pend);
    }
    if (rc != BLT_OK) {  // drain whatever is in flight before the pipe is reused
        cudaStreamSynchronize(pipe->s_h2d);
        cudaStreamSynchronize(pipe->s_comp);
        cudaStreamSynchronize(pipe->s_d2h);
    }
    s->ctx->give_back(std::move(pipe));
    if (rc == BLT_OK) *out_len = off;
    return rc;
}

}  // namespace

int tokenize_host(blt_strategy *s, const uint8_t *in, size_t n, size_t chunk, int content_type, uint8_t *out,
                  size_t out_cap, size_t *out_len) {
    size_t off = 0;
    if (content_type != BLT_CONTENT_NONE) {  // prepend_content_type_token, lib.rs:284-294
        if (out_cap < 2) return fail(BLT_ERR_CAPACITY, "output capacity exceeded");
        const uint16_t t = blt_content_type_token(content_type);
        out[0] = uint8_t(t >> 8);
        out[1] = uint8_t(t & 0xff);
        off = 2;
    }
    *out_len = off;
    if (n == 0) return BLT_OK;  // zero chunks, pipeline.rs:103-105
    if (chunk == 0 || chunk > n) chunk = n;
    if (s->mode == Mode::Passthrough) {  // PassthroughStrategy is a copy (tokenizer.rs:138-144): no kernel
        if (out_cap - off < n) return fail(BLT_ERR_CAPACITY, "output capacity exceeded");
        std::memcpy(out + off, in, n);
        *out_len = off + n;
        return BLT_OK;
    }
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    // The pipeline moves UNITS of several reference chunks: one H2D, one launch (the kernels keep the
    // walls at multiples of `chunk`, which is exactly the concatenation of the per-chunk results) and one
    // D2H per unit.  Fewer, larger PCIe copies; the output bytes are the same.
    // Pageable caller memory (the reference hands in slices of an mmap and takes a fresh Vec<u8>) would make
    // every cudaMemcpyAsync a synchronous copy through the driver's own staging buffer (~3-5 GB/s, serialised
    // across threads).  Such buffers are staged through the pipe's pinned slots with plain (parallel) memcpys.
    const bool stage_in = !is_pinned_host(in), stage_out = !is_pinned_host(out);
    const bool staged = stage_in || stage_out;
    size_t unit_limit = staged ? (size_t(16) << 20) : (size_t(64) << 20);
    if (const char *v = getenv("BLT_UNIT_MB")) unit_limit = size_t(std::max(1, atoi(v))) << 20;  // tuning knob
    size_t per_unit = 1;
    while (per_unit < 64 && chunk * (per_unit * 2) <= unit_limit && chunk * per_unit < n) per_unit *= 2;
    const size_t unit = chunk * per_unit;
    ChunkSource src;
    src.n = n; src.chunk = unit; src.wall = chunk; src.first = 0; src.stride = 1;
    // Copies in and out run side by side, but a unit's output only exists once its whole input has arrived:
    // the D2H stream starts one unit late and trails the H2D stream by one unit at the end.  Both exposed
    // times shrink with the unit, so the schedule ramps up from a single chunk (chunk, 2, 4, ...), runs full
    // units in the middle and ramps down again (..., 4, 2, chunk).
    {
        const size_t n_chunks = (n + chunk - 1) / chunk;
        size_t ramp_chunks = 0;
        for (size_t u = per_unit / 2; u >= 1; u /= 2) ramp_chunks += u;
        const bool ramp = per_unit > 1 && n_chunks >= 4 * per_unit;
        const size_t body_end = ramp ? n_chunks - ramp_chunks : n_chunks;
        size_t c = 0;
        auto push = [&](size_t k) {
            const size_t o = c * chunk;
            src.units.emplace_back(o, std::min(k * chunk, n - o));
            c += k;
        };
        if (ramp) for (size_t u = 1; u < per_unit; u *= 2) push(u);
        while (c < body_end) push(std::min(per_unit, body_end - c));
        if (ramp) for (size_t u = per_unit / 2; u >= 1; u /= 2) push(u);
    }
    src.count = src.units.size();
    return run_host_units(s, src, in, out, out_cap, off, std::min(unit, n), stage_in, stage_out, out_len);
}

// Host memory -> host memory detokenizer: units of 64 MiB of tokens through the same three-stream pipeline
// (any token boundary is a valid cut: tokens expand independently).
int detokenize_host(blt_strategy *s, const uint8_t *in, size_t n_bytes, uint8_t *out, size_t out_cap, size_t *out_len) {
    *out_len = 0;
    if (n_bytes == 0) return BLT_OK;
    if (s->mode == Mode::Passthrough) {
        if (out_cap < n_bytes) return fail(BLT_ERR_CAPACITY, "output capacity exceeded");
        std::memcpy(out, in, n_bytes);
        *out_len = n_bytes;
        return BLT_OK;
    }
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    int rc = s->ensure_detok();
    if (rc) return rc;
    const bool stage_in = !is_pinned_host(in), stage_out = !is_pinned_host(out);
    const size_t unit = std::min(n_bytes, (stage_in || stage_out) ? (size_t(16) << 20) : (size_t(64) << 20));
    ChunkSource src;
    src.n = n_bytes; src.chunk = unit; src.first = 0; src.count = (n_bytes + unit - 1) / unit; src.stride = 1;
    src.detok = true;
    return run_host_units(s, src, in, out, out_cap, 0, unit, stage_in, stage_out, out_len);
}

// ====================================================================================================
// run_tokenizer, file to file
// ====================================================================================================
namespace {

struct OutFile {
    int fd = -1;
    bool seekable = false;
    uint8_t *map = nullptr;  // regular files: the output is mapped at its upper-bound size and trimmed at the end
    size_t map_len = 0;      // (pwrite()s from several threads would serialise on the inode lock)
    int write_at(const uint8_t *p, size_t n, uint64_t off) {  // memcpy into the mapping, else a pwrite loop
        if (map) {
            if (off + n > map_len) return fail(BLT_ERR_CAPACITY, "output exceeds its upper bound");
            std::memcpy(map + off, p, n);
            return BLT_OK;
        }
        while (n) {
            const ssize_t w = seekable ? pwrite(fd, p, n, off_t(off)) : write(fd, p, n);
            if (w < 0) {
                if (errno == EINTR) continue;
                return fail(BLT_ERR_IO, std::string("write failed: ") + std::strerror(errno));
            }
            p += w; n -= size_t(w); off += uint64_t(w);
        }
        return BLT_OK;
    }
};

struct GpuShard {
    int device = 0;
    uint64_t total = 0;  // output bytes this GPU produced
    int rc = BLT_OK;
    std::string err;
};

// The only thing the per-GPU pipelines share: the output length of every chunk.  Chunk k's file
// offset is the sum of the lengths of chunks 0..k-1 (the reference's ordered writer, pipeline.rs:153-168,
// as a prefix); a pipeline that has chunk k's bytes ready waits here until that sum is known.
struct OffsetBoard {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int64_t> len;   // -1 until chunk k's output length is known
    std::vector<uint64_t> cum;  // cum[k] = output bytes of chunks 0..k-1, valid for k <= cursor
    size_t cursor = 0;
    bool failed = false;
    void init(size_t n_chunks) { len.assign(n_chunks, -1); cum.assign(n_chunks + 1, 0); }
    void publish(size_t k, uint64_t n) {
        {
            std::lock_guard<std::mutex> lk(mu);
            len[k] = int64_t(n);
            while (cursor < len.size() && len[cursor] >= 0) { cum[cursor + 1] = cum[cursor] + uint64_t(len[cursor]); ++cursor; }
        }
        cv.notify_all();
    }
    void fail_all() {
        { std::lock_guard<std::mutex> lk(mu); failed = true; }
        cv.notify_all();
    }
    // Blocks until every earlier chunk's length is known; -1 if some pipeline failed.
    int64_t base_of(size_t k) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return failed || cursor >= k; });
        return failed ? -1 : int64_t(cum[k]);
    }
};

// Fresh tmpfs / page-cache pages are the slowest part of writing the output: eight threads storing into a
// newly truncated shared mapping fault pages in at ~8 GB/s together, whereas fallocate() produces them at
// ~14 GB/s from ONE thread and stores into pages that exist run at 20-90 GB/s (tools/tmpfs_probe.cpp, numbers
// in DESIGN.md).  This thread walks ahead of the writers - it starts before the CUDA context is created, so
// the first gigabytes are ready by the time the first chunk comes back - allocating the pages and mapping
// them.  Purely an accelerator: if fallocate is not supported the writers fault pages in as before.
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
class OutPrealloc {
  public:
    enum State { READY = 0, NO_SPACE = 1, UNSUPPORTED = 2 };
    ~OutPrealloc() { finish(); }
    void start(int fd, uint8_t *map, uint64_t bound, uint64_t sure) {
        fd_ = fd; map_ = map; bound_ = bound;
        want_.store(std::min(sure, bound));
        running_ = true;
        th_ = std::thread([this] { loop(); });
    }
    bool running() const { return running_; }
    void enable_mapping() { map_enabled_.store(true, std::memory_order_release); }
    // A writer is about to store [off, off+len) into the mapping: asks for the pages (and `ahead` more) and waits
    // until they exist.  A store into a page of the sparse mapping that cannot be allocated (tmpfs full, quota)
    // would raise SIGBUS and kill the process; the reference returns an io::Error there, and so does this:
    // NO_SPACE -> the caller fails with BLT_ERR_IO; UNSUPPORTED (no fallocate on this filesystem) -> the caller
    // writes that range with pwrite(), which reports ENOSPC itself.
    State ensure(uint64_t off, uint64_t len, uint64_t ahead, int *os_errno) {
        const uint64_t need = std::min(bound_, off + len);
        const uint64_t w = std::min(bound_, off + len + ahead);
        uint64_t cur = want_.load(std::memory_order_relaxed);
        while (cur < w && !want_.compare_exchange_weak(cur, w, std::memory_order_relaxed)) {}
        for (;;) {
            if (done_.load(std::memory_order_acquire) >= need) return READY;
            const int st = state_.load(std::memory_order_acquire);
            if (st != READY) {
                if (allocated_.load(std::memory_order_acquire) >= need) return READY;  // the pages exist, mapped or not
                *os_errno = errno_.load();
                return State(st);
            }
            std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
    }
    void finish() {
        stop_.store(true);
        if (th_.joinable()) th_.join();
    }
    uint64_t prepared() const { return done_.load(std::memory_order_relaxed); }

  private:
    // Two threads, one behind the other: the first allocates (and zeroes) the pages with fallocate (14 GB/s on the pool's
    // hosts), the second maps them into the page table with MADV_POPULATE_WRITE (7.6 GB/s); in one thread the two add up
    // to 4-5 GB/s, which was the file pipeline's ceiling.
    void loop() {
        constexpr uint64_t kPiece = uint64_t(32) << 20;
        std::atomic<uint64_t> &allocated = allocated_;
        std::atomic<bool> alloc_over{false};
        std::thread mapper;
        if (populate_) mapper = std::thread([&] {
            // MADV_POPULATE_WRITE holds the process's mmap lock (shared) for as long as a call lasts, and everything that
            // maps memory needs it exclusively: thread stacks, cudaMalloc, cudaHostAlloc.  So the mapping starts only
            // when the pipelines are set up (enable_mapping); by then the allocation is far ahead and the mapping runs at
            // its full rate (4 MiB pieces were measured: 4 GB/s instead of 7.6).
            constexpr uint64_t kMapPiece = uint64_t(32) << 20;
            uint64_t done = 0;
            while (done < bound_ && !stop_.load(std::memory_order_relaxed)) {
                const uint64_t al = allocated.load(std::memory_order_acquire);
                if (done >= al || !map_enabled_.load(std::memory_order_acquire)) {
                    if (alloc_over.load(std::memory_order_acquire) && done >= allocated.load(std::memory_order_acquire)) break;
                    std::this_thread::sleep_for(std::chrono::microseconds(50));
                    continue;
                }
                const uint64_t len = std::min(kMapPiece, al - done);
                (void)madvise(map_ + done, size_t((len + 4095) & ~uint64_t(4095)), MADV_POPULATE_WRITE);  // best effort
                done += len;
                done_.store(done, std::memory_order_release);
            }
        });
        uint64_t done = 0;
        while (!stop_.load(std::memory_order_relaxed) && done < bound_) {
            const uint64_t w = want_.load(std::memory_order_relaxed);
            if (done >= w) { std::this_thread::sleep_for(std::chrono::microseconds(100)); continue; }
            const uint64_t len = std::min(kPiece, bound_ - done);
            if (fallocate(fd_, 0, off_t(done), off_t(len)) != 0) {
                const int e = errno;
                errno_.store(e);
                state_.store((e == ENOSPC || e == EDQUOT || e == EFBIG) ? NO_SPACE : UNSUPPORTED, std::memory_order_release);
                break;
            }
            done += len;
            allocated.store(done, std::memory_order_release);
            if (!populate_) done_.store(done, std::memory_order_release);  // BLT_NO_POPULATE: the writers take the minor faults
        }
        alloc_over.store(true, std::memory_order_release);
        if (mapper.joinable()) mapper.join();
    }
    int fd_ = -1;
    uint8_t *map_ = nullptr;
    uint64_t bound_ = 0;
    bool running_ = false;
    bool populate_ = getenv("BLT_NO_POPULATE") == nullptr;
    std::atomic<uint64_t> want_{0}, done_{0}, allocated_{0};
    std::atomic<int> state_{READY}, errno_{0};
    std::atomic<bool> stop_{false}, map_enabled_{false};
    std::thread th_;
};

// blt_run_tokenizer is called once per file; in a long-lived process (the Python binding, a service, bench.py) the
// device-side resources of one call are what the next call needs: contexts are kept per device for the life of the
// process and their pipes (pinned staging slots: cudaHostAlloc runs at about 1 GB/s, 144 MiB per pipe) go back to
// the context's pool instead of being freed.  The CLI makes one call and exits, so it frees nothing either.
blt_ctx *file_ctx_for(int device, int *rc_out) {
    static std::mutex mu;
    static std::vector<blt_ctx *> cache;
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() <= size_t(device)) cache.resize(size_t(device) + 1, nullptr);
    if (!cache[size_t(device)]) {
        blt_ctx *c = nullptr;
        *rc_out = blt_ctx_create(device, &c);
        if (*rc_out != BLT_OK) return nullptr;
        cache[size_t(device)] = c;
    }
    *rc_out = BLT_OK;
    return cache[size_t(device)];
}

int build_like(blt_ctx *ctx, const blt_strategy *proto, blt_strategy **out) {
    switch (proto->mode) {
        case Mode::Basic: return blt_strategy_basic(ctx, out);
        case Mode::Passthrough: return blt_strategy_passthrough(ctx, out);
        default: {
            std::vector<uint16_t> l, r, v;
            for (const auto &m : proto->rules) { l.push_back(m.left); r.push_back(m.right); v.push_back(m.value); }
            return blt_strategy_bpe_from_pairs(ctx, l.data(), r.data(), v.data(), l.size(), out);
        }
    }
}

}  // namespace
}  // namespace bltc

using namespace bltc;

namespace {
// BLT_LOG=1 prints stage timings to stderr (the reference logs through `tracing` + RUST_LOG, main.rs:83-85).
struct StageLog {
    bool on = getenv("BLT_LOG") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char *what, int dev = -1) {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (dev >= 0) std::fprintf(stderr, "[blt %9.2f ms] gpu%d %s\n", ms, dev, what);
        else std::fprintf(stderr, "[blt %9.2f ms] %s\n", ms, what);
    }
};
}  // namespace

extern "C" int blt_run_tokenizer(const blt_core_config *cfg) {
    if (!cfg) return fail(BLT_ERR_INVALID_INPUT, "NULL config");
    StageLog slog;
    // ---- CoreConfig::new_from_cli (lib.rs:149-174): threads, chunk size, merges, in this order ----
    const size_t threads = blth::determine_thread_count(cfg->has_threads != 0, cfg->threads);
    bool has_cli_chunk = false;
    size_t cli_chunk = 0;
    if (cfg->chunk_size) {
        const blth::Error e = blth::parse_chunk_size(cfg->chunk_size, &cli_chunk);
        if (e) return fail(BLT_ERR_INVALID_INPUT, e.msg);  // lib.rs:176-182
        has_cli_chunk = true;
    }
    blth::MergeList rules;
    const bool has_merges = cfg->merges_file != nullptr;
    if (has_merges) {
        // a long-lived process tokenizes many files with one merges file: the parsed list is kept while the file's
        // identity (device, inode, size, mtime) does not change (parsing 32 768 lines takes 48 ms)
        static std::mutex cache_mu;
        static struct { dev_t dev = 0; ino_t ino = 0; off_t size = -1; struct timespec mtim{}; blth::MergeList rules; } cache;
        struct stat mst;
        const bool have_stat = stat(cfg->merges_file, &mst) == 0;
        bool hit = false;
        if (have_stat) {
            std::lock_guard<std::mutex> lk(cache_mu);
            if (cache.size == mst.st_size && cache.dev == mst.st_dev && cache.ino == mst.st_ino &&
                cache.mtim.tv_sec == mst.st_mtim.tv_sec && cache.mtim.tv_nsec == mst.st_mtim.tv_nsec) {
                rules = cache.rules;
                hit = true;
            }
        }
        if (!hit) {
            const blth::Error e = blth::load_merges_file(cfg->merges_file, &rules);
            if (e) return fail(BLT_ERR_INVALID_INPUT, "Failed to load BPE merges: " + e.msg);  // lib.rs:194-201
            if (have_stat) {
                std::lock_guard<std::mutex> lk(cache_mu);
                cache.dev = mst.st_dev; cache.ino = mst.st_ino; cache.size = mst.st_size; cache.mtim = mst.st_mtim;
                cache.rules = rules;
            }
        }
    }
    const unsigned memcap = cfg->has_memcap ? cfg->memcap : 80u;  // lib.rs:170
    // ---- run_tokenizer (lib.rs:245-267) ----
    const size_t chunk = blth::effective_chunk_size(has_cli_chunk, cli_chunk, threads, memcap, 0);

    // setup_io (io_handler.rs:51-76): input first, then output
    int in_fd = 0;
    bool in_is_file = cfg->input != nullptr;
    const uint8_t *map = nullptr;
    size_t n = 0;
    if (in_is_file) {
        in_fd = open(cfg->input, O_RDONLY);
        if (in_fd < 0)
            return fail(errno == ENOENT ? BLT_ERR_NOT_FOUND : BLT_ERR_IO,
                        std::string(std::strerror(errno)) + " (os error " + std::to_string(errno) + ")");
        struct stat st;
        if (fstat(in_fd, &st) != 0) { close(in_fd); return fail(BLT_ERR_IO, "fstat failed"); }
        n = size_t(st.st_size);
        if (n) {
            void *p = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, in_fd, 0);
            if (p == MAP_FAILED) { close(in_fd); return fail(BLT_ERR_IO, std::string("mmap failed: ") + std::strerror(errno)); }
            map = static_cast<const uint8_t *>(p);
            madvise(p, n, MADV_SEQUENTIAL);
        }
    }
    // No device, no work: checked BEFORE the output is created and truncated, so that a failure does not clobber an
    // existing output file (passthrough is a host copy and needs no device).
    int n_dev = 0;
    if (!cfg->passthrough) {
        const int drc = blt_device_count(&n_dev);
        if (drc) {
            if (map) munmap(const_cast<uint8_t *>(map), n);
            if (in_is_file) close(in_fd);
            return drc;  // no CPU fallback
        }
    }
    OutFile of;
    if (cfg->output) {
        of.fd = open(cfg->output, O_WRONLY | O_CREAT | O_TRUNC, 0644);  // File::create
        if (of.fd < 0) {
            const int e = errno;
            if (map) munmap(const_cast<uint8_t *>(map), n);
            if (in_is_file) close(in_fd);
            return fail(e == ENOENT ? BLT_ERR_NOT_FOUND : BLT_ERR_IO,
                        std::string(std::strerror(e)) + " (os error " + std::to_string(e) + ")");
        }
        of.seekable = (lseek(of.fd, 0, SEEK_CUR) != (off_t)-1);  // a FIFO or tty cannot be pwritten
    } else {
        of.fd = 1;
        of.seekable = false;
    }
    auto close_io = [&]() {
        if (map) munmap(const_cast<uint8_t *>(map), n);
        if (in_is_file) close(in_fd);
        if (cfg->output) close(of.fd);
    };

    uint64_t prefix = 0;
    if (cfg->content_type != BLT_CONTENT_NONE) {  // prepend_content_type_token (lib.rs:284-294)
        const uint16_t t = blt_content_type_token(cfg->content_type);
        const uint8_t be[2] = {uint8_t(t >> 8), uint8_t(t & 0xff)};
        int rc = of.write_at(be, 2, 0);
        if (rc) { close_io(); return rc; }
        prefix = 2;
    }

    // select_strategy (lib.rs:271-282): passthrough > bpe > basic
    const Mode mode = cfg->passthrough ? Mode::Passthrough : (has_merges ? Mode::BpePairs : Mode::Basic);

    if (mode == Mode::Passthrough) {  // a copy: no device work (tokenizer.rs:138-144)
        int rc = BLT_OK;
        if (in_is_file) {
            if (n) rc = of.write_at(map, n, prefix);
        } else {
            std::vector<uint8_t> buf(chunk);
            uint64_t off = prefix;
            for (;;) {
                const ssize_t r = read(0, buf.data(), buf.size());
                if (r < 0) { if (errno == EINTR) continue; rc = fail(BLT_ERR_IO, "read failed"); break; }
                if (r == 0) break;
                rc = of.write_at(buf.data(), size_t(r), off);
                if (rc) break;
                off += uint64_t(r);
            }
        }
        close_io();
        return rc;
    }

    // ---- the mmap path's output: mapped at its upper bound, pages prepared in the background from now on
    // (the CUDA context creation below takes 0.4-1.5 s, which is when the first gigabytes get ready) ----
    const size_t n_chunks = (in_is_file && n) ? (n + chunk - 1) / chunk : 0;
    OutPrealloc pre;
    if (n_chunks && of.seekable && getenv("BLT_NO_MMAP_OUT") == nullptr) {
        struct stat ost;
        const size_t bound = size_t(prefix) + 2 * n;
        if (fstat(of.fd, &ost) == 0 && S_ISREG(ost.st_mode) && ftruncate(of.fd, off_t(bound)) == 0) {
            // O_WRONLY descriptors cannot be mapped shared: reopen read-write through /proc
            const std::string self = "/proc/self/fd/" + std::to_string(of.fd);
            const int rw = open(self.c_str(), O_RDWR);
            if (rw >= 0) {
                void *m = mmap(nullptr, bound, PROT_READ | PROT_WRITE, MAP_SHARED, rw, 0);
                close(rw);
                if (m != MAP_FAILED) { of.map = static_cast<uint8_t *>(m); of.map_len = bound; }
            }
            if (!of.map) (void)!ftruncate(of.fd, off_t(prefix));
        }
        // every strategy but passthrough writes at least n bytes (BPE: >= n/2 tokens of 2 bytes; basic: 2n)
        if (of.map && getenv("BLT_NO_PREALLOC") == nullptr)
            pre.start(of.fd, of.map, bound, mode == Mode::Basic ? bound : prefix + n);
    }
    auto cleanup = [&]() {
        pre.finish();
        if (of.map) {  // an early error: give the mapping back and leave only what was written
            munmap(of.map, of.map_len);
            of.map = nullptr;
            (void)!ftruncate(of.fd, off_t(prefix));
        }
        close_io();
    };

    slog.mark("config parsed, io open");
    int rc = BLT_OK;
    // default: one GPU.  The host side (page cache, PCIe) bounds file-to-file long before one B200 does, and
    // every further context costs about a second of start-up (host pipeline).
    int n_gpus = cfg->num_gpus > 0 ? std::min(cfg->num_gpus, n_dev) : 1;
    if (!of.seekable) n_gpus = 1;  // a stream can only be written front to back

    if (!in_is_file) {
        // ---- stdin: run_stream_pipeline (pipeline.rs:196-240).  The reference cuts a chunk at
        // whatever one read() returns (pipeline.rs:310-318), which is not reproducible for BPE; here
        // every chunk is filled to C bytes before it is processed (documented divergence;
        // identical for Basic, whose output does not depend on the cut).
        blt_ctx *ctx = nullptr;
        rc = blt_ctx_create(0, &ctx);
        if (rc) { cleanup(); return rc; }
        blt_strategy *st = nullptr;
        rc = has_merges ? [&] {
            std::vector<uint16_t> l, r, v;
            for (const auto &m : rules) { l.push_back(m.left); r.push_back(m.right); v.push_back(m.value); }
            return blt_strategy_bpe_from_pairs(ctx, l.data(), r.data(), v.data(), l.size(), &st);
        }() : blt_strategy_basic(ctx, &st);
        std::vector<uint8_t> ibuf(chunk), obuf(2 * chunk);
        uint64_t off = prefix;
        while (rc == BLT_OK) {
            size_t got = 0;
            while (got < chunk) {
                const ssize_t r = read(0, ibuf.data() + got, chunk - got);
                if (r < 0) { if (errno == EINTR) continue; rc = fail(BLT_ERR_IO, "read failed"); break; }
                if (r == 0) break;
                got += size_t(r);
            }
            if (rc || got == 0) break;
            size_t olen = 0;
            rc = blt_process_chunk(st, ibuf.data(), got, obuf.data(), obuf.size(), &olen);
            if (rc == BLT_OK) rc = of.write_at(obuf.data(), olen, off);
            off += olen;
            if (got < chunk) break;
        }
        if (st) blt_strategy_destroy(st);
        blt_ctx_destroy(ctx);
        cleanup();
        return rc;
    }

    // ---- mmap path: run_mmap_pipeline (pipeline.rs:56-131) over n_gpus devices ----
    if (n_chunks == 0) { cleanup(); return BLT_OK; }  // empty file -> empty output (pipeline.rs:103-105)
    if (size_t(n_gpus) > n_chunks) n_gpus = int(n_chunks);
    // Chunk k goes to GPU k % G: the pipelines advance through the file together, so a chunk's offset (the
    // lengths of all earlier chunks) is known almost as soon as its own bytes are back on the host.
    std::vector<GpuShard> shards(static_cast<size_t>(n_gpus));
    OffsetBoard board;
    board.init(n_chunks);
    blt_strategy proto;
    proto.mode = mode;
    proto.rules = rules;

    std::mutex drain_mu;
    std::condition_variable drain_cv;
    int drained = 0;
    auto worker = [&](size_t g) {
        GpuShard &sh = shards[g];
        blt_ctx *ctx = nullptr;
        blt_strategy *st = nullptr;
        std::unique_ptr<Pipe> pipe;
        int rc = BLT_OK;
        ctx = file_ctx_for(sh.device, &rc);
        if (rc == BLT_OK && cudaSetDevice(sh.device) != cudaSuccess) rc = fail(BLT_ERR_CUDA, "cudaSetDevice failed");
        slog.mark("context ready", sh.device);
        if (rc == BLT_OK) rc = build_like(ctx, &proto, &st);
        slog.mark("strategy built", sh.device);
        // host copies are split over `--threads / gpus` threads (at most 8), like the reference's workers: one half
        // stages input (this thread + helpers), the other half drains output (the drainer thread + helpers), so that
        // the two directions run side by side
        const size_t host_threads = std::min<size_t>(8, std::max<size_t>(2, threads / size_t(n_gpus)));
        HostPool io_in((host_threads + 1) / 2 - 1), io_out(std::max<size_t>(1, host_threads / 2) - 1);
        constexpr size_t kPiece = size_t(2) << 20;
        auto par_copy = [&](HostPool &io, uint8_t *dst, const uint8_t *src_p, size_t len) {
            const size_t parts = std::min(io.width(), (len + kPiece - 1) / kPiece);
            if (parts <= 1) { std::memcpy(dst, src_p, len); return; }
            const size_t per = ((len + parts - 1) / parts + 4095) & ~size_t(4095);
            io.run(parts, [&](size_t i) {
                const size_t lo = i * per;
                if (lo < len) std::memcpy(dst + lo, src_p + lo, std::min(per, len - lo));
            });
        };
        auto par_memcpy = [&](uint8_t *dst, const uint8_t *src_p, size_t len) { par_copy(io_out, dst, src_p, len); };
        auto put = [&](const uint8_t *buf, size_t len, uint64_t off) -> int {
            if (of.map) {  // mapped output: plain stores, split over the helper threads, into pages that exist
                if (off + len > of.map_len) return fail(BLT_ERR_CAPACITY, "output exceeds its upper bound");
                int e = 0;
                OutPrealloc::State st = OutPrealloc::UNSUPPORTED;
                if (pre.running()) {
                    st = pre.ensure(off, len, uint64_t(256) << 20, &e);
                } else if (fallocate(of.fd, 0, off_t(off), off_t(len)) == 0) {  // BLT_NO_PREALLOC: the writer allocates its own range
                    st = OutPrealloc::READY;
                } else {
                    e = errno;
                    if (e == ENOSPC || e == EDQUOT || e == EFBIG) st = OutPrealloc::NO_SPACE;
                }
                if (st == OutPrealloc::NO_SPACE)
                    return fail(BLT_ERR_IO, std::string(std::strerror(e)) + " (os error " + std::to_string(e) + ")");
                if (st == OutPrealloc::READY) {
                    par_memcpy(of.map + off, buf, len);
                    return BLT_OK;
                }
                // no fallocate on this filesystem: pwrite reports a full device as an error instead of a SIGBUS
                while (len) {
                    const ssize_t w = pwrite(of.fd, buf, len, off_t(off));
                    if (w < 0) {
                        if (errno == EINTR) continue;
                        return fail(BLT_ERR_IO, std::string(std::strerror(errno)) + " (os error " + std::to_string(errno) + ")");
                    }
                    buf += w; len -= size_t(w); off += uint64_t(w);
                }
                return BLT_OK;
            }
            return of.write_at(buf, len, off);
        };
        uint64_t produced = 0;
        double t_in = 0, t_out = 0, t_wait = 0;  // host seconds: staging copies in, copies out, waiting for D2H
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(now() - a).count(); };
        if (rc == BLT_OK) {
            ChunkSource src;
            // the chunks k with blt_file_chunk_device(k, n_gpus) == g
            src.n = n; src.chunk = chunk; src.first = g; src.stride = size_t(n_gpus);
            src.count = (n_chunks > g) ? (n_chunks - g + size_t(n_gpus) - 1) / size_t(n_gpus) : 0;
            pipe = ctx->acquire();
            rc = pipe->ensure(std::min(chunk, n), std::min(kSlots, std::max<size_t>(src.count, 1)), true);
            slog.mark("pipe buffers allocated", sh.device);
            // Output is drained by a second thread: it waits for chunk k's D2H, asks the board for the chunk's offset
            // and copies pinned -> output mapping while this thread stages the input of the chunks behind it.
            struct DrainItem { Slot *sl; size_t len, id; };
            std::mutex dmu;
            std::condition_variable dcv;
            std::deque<DrainItem> dq;
            bool dstop = false;
            int drc = BLT_OK;
            std::string derr;
            std::vector<size_t> slot_submitted(kSlots, 0), slot_drained(kSlots, 0);
            std::thread drainer([&] {
                cudaSetDevice(sh.device);
                for (;;) {
                    DrainItem it;
                    {
                        std::unique_lock<std::mutex> lk(dmu);
                        dcv.wait(lk, [&] { return dstop || !dq.empty(); });
                        if (dq.empty()) return;
                        it = dq.front();
                        dq.pop_front();
                    }
                    int w = BLT_OK;
                    if (drc == BLT_OK) {
                        auto t0 = now();
                        if (cudaEventSynchronize(it.sl->ev_d2h) != cudaSuccess) w = fail(BLT_ERR_CUDA, "cudaEventSynchronize failed");
                        // Basic is fixed-ratio: the offset is known up front; otherwise ask the board
                        const int64_t base = (w != BLT_OK) ? -1 : (mode == Mode::Basic) ? int64_t(2 * it.id * chunk) : board.base_of(it.id);
                        if (w == BLT_OK && base < 0) w = fail(BLT_ERR_IO, "another GPU pipeline failed");
                        t_wait += since(t0);
                        if (w == BLT_OK) {
                            t0 = now();
                            w = put(it.sl->h_out, it.len, prefix + uint64_t(base));
                            t_out += since(t0);
                            produced += it.len;
                        }
                    }
                    {
                        std::lock_guard<std::mutex> lk(dmu);
                        if (w != BLT_OK && drc == BLT_OK) { drc = w; derr = blt_last_error(); board.fail_all(); }
                        ++slot_drained[size_t(it.sl - pipe->slots.data())];
                    }
                    dcv.notify_all();
                }
            });
            pre.enable_mapping();  // this pipeline's threads and buffers exist: the page-table population may start
            if (rc == BLT_OK && src.count) {
                rc = run_slots(
                    st, *pipe, src,
                    [&](size_t id, Slot &sl) {
                        const auto t0 = now();
                        par_copy(io_in, sl.h_in, map + id * chunk, src.len_of(id));  // page cache -> pinned
                        t_in += since(t0);
                        return static_cast<const uint8_t *>(sl.h_in);
                    },
                    [&](size_t id, Slot &sl, size_t len) -> int {
                        board.publish(id, len);  // the length is known before the bytes are back
                        const size_t si = size_t(&sl - pipe->slots.data());
                        {   // the slot's pinned output buffer is free once its previous chunk has been drained
                            std::unique_lock<std::mutex> lk(dmu);
                            dcv.wait(lk, [&] { return drc != BLT_OK || slot_drained[si] == slot_submitted[si]; });
                            if (drc != BLT_OK) return fail(drc, derr);
                        }
                        CUDA_TRY(cudaMemcpyAsync(sl.h_out, sl.d_out, len, cudaMemcpyDeviceToHost, pipe->s_d2h));
                        CUDA_TRY(cudaEventRecord(sl.ev_d2h, pipe->s_d2h));
                        {
                            std::lock_guard<std::mutex> lk(dmu);
                            ++slot_submitted[si];
                            dq.push_back(DrainItem{&sl, len, id});
                        }
                        dcv.notify_all();
                        return BLT_OK;
                    });
            }
            {
                std::lock_guard<std::mutex> lk(dmu);
                dstop = true;
            }
            dcv.notify_all();
            drainer.join();
            if (rc == BLT_OK && drc != BLT_OK) rc = fail(drc, derr);
        }
        sh.total = produced;
        if (rc != BLT_OK) { sh.rc = rc; sh.err = blt_last_error(); board.fail_all(); }
        slog.mark("pipeline drained", sh.device);
        if (slog.on)
            std::fprintf(stderr, "[blt] gpu%d host seconds: copy-in %.3f, copy-out %.3f, drainer waiting for the device / the offset %.3f (%zu + %zu threads)\n",
                         sh.device, t_in, t_out, t_wait, io_in.width(), io_out.width());
        {   // from here on this thread only gives device resources back; the files are the main thread's
            std::lock_guard<std::mutex> lk(drain_mu);
            ++drained;
        }
        drain_cv.notify_all();
        if (pipe) {
            if (rc != BLT_OK) {  // drain whatever is in flight before the pipe is reused
                cudaStreamSynchronize(pipe->s_h2d);
                cudaStreamSynchronize(pipe->s_comp);
                cudaStreamSynchronize(pipe->s_d2h);
            }
            ctx->give_back(std::move(pipe));
        }
        if (st) blt_strategy_destroy(st);
        slog.mark("device resources returned to the pool", sh.device);
    };

    std::vector<std::thread> pool;
    for (size_t g = 0; g < size_t(n_gpus); ++g) {
        shards[g].device = int(g);
        pool.emplace_back(worker, g);
    }
    {
        std::unique_lock<std::mutex> lk(drain_mu);
        drain_cv.wait(lk, [&] { return drained == n_gpus; });
    }
    // Every chunk is in the output mapping.  Unmapping gigabytes of populated pages takes 0.1-0.3 s, the same
    // order as freeing the pinned buffers and the context, so the two teardowns run side by side.
    pre.finish();
    if (slog.on) std::fprintf(stderr, "[blt] output pages prepared ahead of the writers: %.2f GiB\n", double(pre.prepared()) / double(1 << 30));
    rc = BLT_OK;
    for (const auto &sh : shards)
        if (sh.rc != BLT_OK && rc == BLT_OK) { rc = sh.rc; fail(sh.rc, sh.err); }  // first error in chunk order
    // Tearing down gigabytes of populated mappings takes 50-70 ms per GiB and nobody waits for it: the bytes are in the
    // page cache since they were stored.  The output is trimmed here; the two munmaps run on a detached thread.
    {
        uint8_t *om = of.map;
        const size_t om_len = of.map_len;
        if (of.map) {  // trim the mapping's upper bound down to what was produced
            uint64_t total = prefix;
            for (const auto &sh : shards) total += sh.total;
            of.map = nullptr;
            if (ftruncate(of.fd, off_t(total)) != 0 && rc == BLT_OK) rc = fail(BLT_ERR_IO, "ftruncate failed");
            slog.mark("output trimmed");
        }
        uint8_t *im = const_cast<uint8_t *>(map);
        const size_t im_len = n;
        map = nullptr;
        if (om || im) std::thread([om, om_len, im, im_len] {
            if (om) munmap(om, om_len);
            if (im) munmap(im, im_len);
        }).detach();
    }
    slog.mark("mappings handed to the teardown thread");
    for (auto &t : pool) t.join();
    slog.mark("all shards done");
    cleanup();
    return rc;
}

// ---- detokenizer, file to file ----------------------------------------------------------------------
extern "C" int blt_run_detokenizer(const blt_core_config *cfg) {
    if (!cfg) return fail(BLT_ERR_INVALID_INPUT, "NULL config");
    blth::MergeList rules;
    if (cfg->merges_file) {
        const blth::Error e = blth::load_merges_file(cfg->merges_file, &rules);
        if (e) return fail(BLT_ERR_INVALID_INPUT, "Failed to load BPE merges: " + e.msg);
    }
    // input: the whole token stream (mmap, or stdin read to the end)
    std::vector<uint8_t> held;
    const uint8_t *in = nullptr;
    size_t n = 0;
    int in_fd = -1;
    void *map = nullptr;
    if (cfg->input) {
        in_fd = open(cfg->input, O_RDONLY);
        if (in_fd < 0)
            return fail(errno == ENOENT ? BLT_ERR_NOT_FOUND : BLT_ERR_IO,
                        std::string(std::strerror(errno)) + " (os error " + std::to_string(errno) + ")");
        struct stat st;
        if (fstat(in_fd, &st) != 0) { close(in_fd); return fail(BLT_ERR_IO, "fstat failed"); }
        n = size_t(st.st_size);
        if (n) {
            map = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, in_fd, 0);
            if (map == MAP_FAILED) { close(in_fd); return fail(BLT_ERR_IO, std::string("mmap failed: ") + std::strerror(errno)); }
            in = static_cast<const uint8_t *>(map);
        }
    } else {
        uint8_t buf[1 << 16];
        for (;;) {
            const ssize_t r = read(0, buf, sizeof buf);
            if (r < 0) { if (errno == EINTR) continue; return fail(BLT_ERR_IO, "read failed"); }
            if (r == 0) break;
            held.insert(held.end(), buf, buf + r);
        }
        in = held.data();
        n = held.size();
    }
    auto close_in = [&] { if (map) munmap(map, n); if (in_fd >= 0) close(in_fd); };
    OutFile of;
    if (cfg->output) {
        of.fd = open(cfg->output, O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (of.fd < 0) {
            const int e = errno;
            close_in();
            return fail(e == ENOENT ? BLT_ERR_NOT_FOUND : BLT_ERR_IO, std::string(std::strerror(e)) + " (os error " + std::to_string(e) + ")");
        }
        of.seekable = (lseek(of.fd, 0, SEEK_CUR) != (off_t)-1);
    } else {
        of.fd = 1;
    }
    blt_ctx *ctx = nullptr;
    blt_strategy *st = nullptr;
    int rc = blt_ctx_create(0, &ctx);
    if (rc == BLT_OK) {
        if (cfg->passthrough) rc = blt_strategy_passthrough(ctx, &st);
        else if (cfg->merges_file) {
            std::vector<uint16_t> l, r, v;
            for (const auto &m : rules) { l.push_back(m.left); r.push_back(m.right); v.push_back(m.value); }
            rc = blt_strategy_bpe_from_pairs(ctx, l.data(), r.data(), v.data(), l.size(), &st);
        } else rc = blt_strategy_basic(ctx, &st);
    }
    std::vector<uint8_t> out(std::max<size_t>(n, 1));
    size_t out_len = 0;
    if (rc == BLT_OK) rc = blt_detokenize_host(st, in, n, cfg->content_type != BLT_CONTENT_NONE, out.data(), out.size(), &out_len);
    if (rc == BLT_OK) rc = of.write_at(out.data(), out_len, 0);
    const std::string err = rc != BLT_OK ? std::string(blt_last_error()) : std::string();
    if (st) blt_strategy_destroy(st);
    if (ctx) blt_ctx_destroy(ctx);
    if (cfg->output) close(of.fd);
    close_in();
    return rc == BLT_OK ? BLT_OK : fail(rc, err);
}
