// pipeline.cu -- host side of the chunk pipeline: the in-memory form (blt_tokenize_host) and the
// file-to-file form (blt_run_tokenizer) of run_mmap_pipeline / run_stream_pipeline
// (blt_core/src/pipeline.rs:56-240).  Chunks are cut at fixed offsets k*C from the start of the input
// (pipeline.rs:73-81), flow through S slots (H2D copy stream -> compute stream -> D2H copy stream)
// and are written strictly in chunk order (pipeline.rs:153-168).  With several GPUs each device gets
// a contiguous range of chunks and its own pipeline; the only cross-GPU datum is each range's output
// length (a host-side prefix), so nothing is exchanged between devices.
#include "pipeline.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
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
    const bool regrow = chunk_cap > cap || (want_pinned && !pinned);
    if (regrow) {
        for (Slot &sl : slots) {
            if (sl.d_in) cudaFree(sl.d_in);
            if (sl.d_out) cudaFree(sl.d_out);
            if (sl.h_in) cudaFreeHost(sl.h_in);
            if (sl.h_out) cudaFreeHost(sl.h_out);
            sl.d_in = sl.d_out = sl.h_in = sl.h_out = nullptr;
        }
        cap = std::max(chunk_cap, cap);
        pinned = pinned || want_pinned;
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
        if (pinned && !sl.h_in) {
            CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_in), cap, cudaHostAllocDefault));
            CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_out), 2 * cap, cudaHostAllocDefault));
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
    if (s_h2d) cudaStreamDestroy(s_h2d);
    if (s_comp) cudaStreamDestroy(s_comp);
    if (s_d2h) cudaStreamDestroy(s_d2h);
    s_h2d = s_comp = s_d2h = nullptr;
    ws.release();
    cap = 0;
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

// The slot-pipelined loop shared by the in-memory and the file pipelines.
//   fetch(k, slot)      -> host pointer to chunk k's input bytes (may stage into slot.h_in)
//   deliver(k, slot, n) -> consume chunk k's n output bytes, which the D2H stream is writing to
//                          `dst`; called after the copy has been enqueued; returns where chunk k's
//                          bytes must land (see users)
// Chunk order is preserved by construction: completion is processed for k = 0,1,2,... in order.
struct ChunkSource {
    size_t n = 0, chunk = 0, first = 0, last = 0;  // units [first,last) of `chunk` bytes of an n-byte input
    size_t wall = 0;                                 // the reference's chunk size inside a unit (0: unit == chunk)
    size_t len_of(size_t k) const { return std::min(chunk, n - k * chunk); }
};

template <class Fetch, class Sink>
int run_slots(blt_strategy *s, Pipe &pipe, const ChunkSource &src, Fetch fetch, Sink sink) {
    const size_t S = std::min(kSlots, src.last - src.first);
    auto issue = [&](size_t k) -> int {
        Slot &sl = pipe.slots[(k - src.first) % S];
        sl.in_len = src.len_of(k);
        const bool reuse = (k - src.first) >= S;
        if (reuse) CUDA_TRY(cudaStreamWaitEvent(pipe.s_h2d, sl.ev_done, 0));   // d_in free once its kernel ran
        const uint8_t *h = fetch(k, sl);
        CUDA_TRY(cudaMemcpyAsync(sl.d_in, h, sl.in_len, cudaMemcpyHostToDevice, pipe.s_h2d));
        CUDA_TRY(cudaEventRecord(sl.ev_h2d, pipe.s_h2d));
        CUDA_TRY(cudaStreamWaitEvent(pipe.s_comp, sl.ev_h2d, 0));
        if (reuse) CUDA_TRY(cudaStreamWaitEvent(pipe.s_comp, sl.ev_d2h, 0));  // d_out free once copied out
        int rc = run_device(s, pipe.ws, sl.d_in, sl.in_len, src.wall, sl.d_out, 2 * pipe.cap, nullptr, pipe.s_comp, &sl.res);
        if (rc) return rc;
        if (sl.res.kind == DeviceResult::IN_SCRATCH)
            CUDA_TRY(cudaMemcpyAsync(sl.h_ctrl, pipe.ws.scratch.ctrl, 32, cudaMemcpyDeviceToHost, pipe.s_comp));
        CUDA_TRY(cudaEventRecord(sl.ev_done, pipe.s_comp));
        return BLT_OK;
    };
    for (size_t k = src.first; k < src.first + S; ++k) {
        int rc = issue(k);
        if (rc) return rc;
    }
    for (size_t k = src.first; k < src.last; ++k) {
        Slot &sl = pipe.slots[(k - src.first) % S];
        CUDA_TRY(cudaEventSynchronize(sl.ev_done));
        if (sl.res.kind == DeviceResult::IN_SCRATCH) {
            int rc = decode_ctrl(sl.h_ctrl, &sl.res);
            if (rc) return rc;
        }
        int rc = sink(k, sl, sl.res.len);  // enqueues the D2H copy on s_d2h and records ev_d2h
        if (rc) return rc;
        if (k + S < src.last) {
            rc = issue(k + S);
            if (rc) return rc;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(pipe.s_d2h));
    return BLT_OK;
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
    size_t per_unit = 1;
    while (per_unit < 8 && chunk * (per_unit * 2) <= (size_t(64) << 20) && chunk * per_unit < n) per_unit *= 2;
    const size_t unit = chunk * per_unit;
    ChunkSource src;
    src.n = n; src.chunk = unit; src.wall = chunk; src.first = 0; src.last = (n + unit - 1) / unit;
    auto pipe = s->ctx->acquire();
    int rc = pipe->ensure(std::min(unit, n), std::min(kSlots, src.last), false);
    if (rc == BLT_OK) {
        rc = run_slots(
            s, *pipe, src, [&](size_t k, Slot &) { return in + k * unit; },
            [&](size_t, Slot &sl, size_t len) -> int {
                if (off + len > out_cap) return fail(BLT_ERR_CAPACITY, "output capacity exceeded");
                CUDA_TRY(cudaMemcpyAsync(out + off, sl.d_out, len, cudaMemcpyDeviceToHost, pipe->s_d2h));
                CUDA_TRY(cudaEventRecord(sl.ev_d2h, pipe->s_d2h));
                off += len;
                return BLT_OK;
            });
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

// ====================================================================================================
// run_tokenizer, file to file
// ====================================================================================================
namespace {

struct OutFile {
    int fd = -1;
    bool seekable = false;
    int write_at(const uint8_t *p, size_t n, uint64_t off) {  // pwrite loop
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
    size_t first = 0, last = 0;      // chunk range
    std::vector<uint8_t> held;       // output produced before this shard's file offset was known
    uint64_t total = 0;
    int rc = BLT_OK;
    std::string err;
};

// Shared between the per-GPU threads: totals of finished shards -> file offsets of later ones.
struct OffsetBoard {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int64_t> total;  // -1 until shard g has produced all its output
    bool failed = false;
    // Offset of shard g relative to shard 0, or -1 if an earlier shard is still running.
    int64_t try_base(size_t g) {
        std::lock_guard<std::mutex> lk(mu);
        int64_t b = 0;
        for (size_t i = 0; i < g; ++i) { if (total[i] < 0) return -1; b += total[i]; }
        return b;
    }
    int64_t wait_base(size_t g) {
        std::unique_lock<std::mutex> lk(mu);
        int64_t b = 0;
        cv.wait(lk, [&] {
            if (failed) return true;
            b = 0;
            for (size_t i = 0; i < g; ++i) { if (total[i] < 0) return false; b += total[i]; }
            return true;
        });
        return failed ? -1 : b;
    }
    void publish(size_t g, int64_t t, bool ok) {
        { std::lock_guard<std::mutex> lk(mu); total[g] = t; if (!ok) failed = true; }
        cv.notify_all();
    }
};

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
        const blth::Error e = blth::load_merges_file(cfg->merges_file, &rules);
        if (e) return fail(BLT_ERR_INVALID_INPUT, "Failed to load BPE merges: " + e.msg);  // lib.rs:194-201
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
    auto cleanup = [&]() {
        if (map) munmap(const_cast<uint8_t *>(map), n);
        if (in_is_file) close(in_fd);
        if (cfg->output) close(of.fd);
    };

    uint64_t prefix = 0;
    if (cfg->content_type != BLT_CONTENT_NONE) {  // prepend_content_type_token (lib.rs:284-294)
        const uint16_t t = blt_content_type_token(cfg->content_type);
        const uint8_t be[2] = {uint8_t(t >> 8), uint8_t(t & 0xff)};
        int rc = of.write_at(be, 2, 0);
        if (rc) { cleanup(); return rc; }
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
        cleanup();
        return rc;
    }

    slog.mark("config parsed, io open");
    int n_dev = 0;
    int rc = blt_device_count(&n_dev);
    slog.mark("device count");
    if (rc) { cleanup(); return rc; }  // no CPU fallback
    int n_gpus = cfg->num_gpus > 0 ? std::min(cfg->num_gpus, n_dev) : n_dev;
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
    const size_t n_chunks = n ? (n + chunk - 1) / chunk : 0;
    if (n_chunks == 0) { cleanup(); return BLT_OK; }  // empty file -> empty output (pipeline.rs:103-105)
    if (size_t(n_gpus) > n_chunks) n_gpus = int(n_chunks);
    std::vector<size_t> bounds(size_t(n_gpus) + 1);
    blt_shard_chunks(n_chunks, n_gpus, bounds.data());
    std::vector<GpuShard> shards(static_cast<size_t>(n_gpus));
    OffsetBoard board;
    board.total.assign(size_t(n_gpus), -1);
    blt_strategy proto;
    proto.mode = mode;
    proto.rules = rules;
    if (mode == Mode::BpePairs)
        for (const auto &r : rules)
            if (r.value < 256) proto.mode = Mode::BpeGeneral;  // cannot happen for a merges.txt; kept for symmetry

    auto worker = [&](size_t g) {
        GpuShard &sh = shards[g];
        blt_ctx *ctx = nullptr;
        blt_strategy *st = nullptr;
        std::unique_ptr<Pipe> pipe;
        int rc = blt_ctx_create(sh.device, &ctx);
        slog.mark("context created", sh.device);
        if (rc == BLT_OK) rc = build_like(ctx, &proto, &st);
        slog.mark("strategy built", sh.device);
        // Fixed-ratio output (Basic): every shard knows its offset up front and streams with pwrite.
        int64_t base = (mode == Mode::Basic) ? int64_t(2 * sh.first * chunk) : (g == 0 ? 0 : -1);
        uint64_t produced = 0;
        if (rc == BLT_OK) {
            ChunkSource src;
            src.n = n; src.chunk = chunk; src.first = sh.first; src.last = sh.last;
            pipe = ctx->acquire();
            rc = pipe->ensure(std::min(chunk, n), std::min(kSlots, src.last - src.first), true);
            slog.mark("pipe buffers allocated", sh.device);
            // Output is delivered one chunk late so chunk k's D2H overlaps chunk k+1's host-side input staging.
            struct Pending { Slot *sl = nullptr; size_t len = 0; } pend;
            auto flush = [&](Pending &p) -> int {
                if (!p.sl) return BLT_OK;
                CUDA_TRY(cudaEventSynchronize(p.sl->ev_d2h));
                if (base < 0) base = board.try_base(g);
                int w = BLT_OK;
                if (base >= 0) {
                    if (!sh.held.empty()) {  // offset just became known: drain what was held back
                        w = of.write_at(sh.held.data(), sh.held.size(), prefix + uint64_t(base));
                        std::vector<uint8_t>().swap(sh.held);
                    }
                    if (w == BLT_OK) w = of.write_at(p.sl->h_out, p.len, prefix + uint64_t(base) + produced);
                } else {
                    sh.held.insert(sh.held.end(), p.sl->h_out, p.sl->h_out + p.len);
                }
                produced += p.len;
                p.sl = nullptr;
                return w;
            };
            if (rc == BLT_OK) {
                rc = run_slots(
                    st, *pipe, src,
                    [&](size_t k, Slot &sl) {
                        std::memcpy(sl.h_in, map + k * chunk, src.len_of(k));  // page cache -> pinned
                        return static_cast<const uint8_t *>(sl.h_in);
                    },
                    [&](size_t, Slot &sl, size_t len) -> int {
                        int w = flush(pend);  // previous chunk: its copy has had a full stage to finish
                        if (w) return w;
                        CUDA_TRY(cudaMemcpyAsync(sl.h_out, sl.d_out, len, cudaMemcpyDeviceToHost, pipe->s_d2h));
                        CUDA_TRY(cudaEventRecord(sl.ev_d2h, pipe->s_d2h));
                        pend.sl = &sl;
                        pend.len = len;
                        return BLT_OK;
                    });
                if (rc == BLT_OK) rc = flush(pend);
            }
        }
        sh.total = produced;
        slog.mark("pipeline drained", sh.device);
        if (rc != BLT_OK) { sh.rc = rc; sh.err = blt_last_error(); }
        board.publish(g, int64_t(produced), rc == BLT_OK);
        if (rc == BLT_OK && !sh.held.empty()) {  // wait for the shards before us, then write our range
            const int64_t b = board.wait_base(g);
            if (b >= 0) {
                rc = of.write_at(sh.held.data(), sh.held.size(), prefix + uint64_t(b));
                if (rc) { sh.rc = rc; sh.err = blt_last_error(); }
            }
        }
        slog.mark("held output written", sh.device);
        if (pipe) { pipe->release(); }
        if (st) blt_strategy_destroy(st);
        if (ctx) blt_ctx_destroy(ctx);
        slog.mark("released", sh.device);
    };

    std::vector<std::thread> pool;
    for (size_t g = 0; g < size_t(n_gpus); ++g) {
        shards[g].device = int(g);
        shards[g].first = bounds[g];
        shards[g].last = bounds[g + 1];
        pool.emplace_back(worker, g);
    }
    for (auto &t : pool) t.join();
    slog.mark("all shards done");
    rc = BLT_OK;
    for (const auto &sh : shards)
        if (sh.rc != BLT_OK && rc == BLT_OK) { rc = sh.rc; fail(sh.rc, sh.err); }  // first error in chunk order
    cleanup();
    return rc;
}
