// cabi.cu -- the extern "C" surface of libblt_cuda.so (include/blt_cuda.h): contexts, strategies,
// the per-chunk / pipelined / device-resident entry points.  No CPU compute path exists here: every
// tokenizing call ends in a kernel launch from kernels.cu or fails.
#include "../../include/blt_cuda.h"
#include "host_config.h"
#include "kernels.cuh"
#include "pipeline.h"

#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#define BLT_VERSION_STRING "0.2.2"  // CARGO_PKG_VERSION of the reference this build mirrors

namespace bltc {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return bltc::fail(e__ == cudaErrorMemoryAllocation ? BLT_ERR_NOMEM : BLT_ERR_CUDA,     \
                              std::string(#expr) + ": " + cudaGetErrorString(e__));                \
    } while (0)

// ---- Workspace: device scratch for one in-flight tokenization -------------------------------------
int Workspace::ensure_scratch(size_t n_elems) {
    if (n_elems <= scratch_elems && d_scratch) return BLT_OK;
    if (d_scratch) cudaFree(d_scratch);
    d_scratch = nullptr;
    scratch_elems = 0;
    const size_t want = std::max<size_t>(n_elems, 1u << 20);
    CUDA_TRY(cudaMalloc(&d_scratch, bltk::sweep_scratch_bytes(want)));
    scratch_elems = want;
    scratch = bltk::sweep_scratch_carve(d_scratch, want);
    return BLT_OK;
}
int Workspace::ensure_work(size_t chunk_bytes) {
    if (!h_ctrl) CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&h_ctrl), 64, cudaHostAllocDefault));
    if (chunk_bytes <= work_chunk) return BLT_OK;
    for (auto &p : d_work) { if (p) cudaFree(p); p = nullptr; }
    if (d_align) cudaFree(d_align);
    d_align = nullptr;
    work_chunk = 0;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&d_work[0]), 2 * chunk_bytes + 64));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&d_work[1]), 2 * chunk_bytes + 64));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&d_align), chunk_bytes + 64));
    work_chunk = chunk_bytes;
    return BLT_OK;
}
void Workspace::release_lanes() {
    for (GenLane &l : lanes) {
        if (l.d_scratch) cudaFree(l.d_scratch);
        for (auto &p : l.d_work) if (p) cudaFree(p);
        if (l.d_align) cudaFree(l.d_align);
        if (l.h_ctrl) cudaFreeHost(l.h_ctrl);
    }
    lanes.clear();
    lane_chunk = 0;
}
int Workspace::ensure_lanes(size_t chunk_bytes, size_t n_lanes) {
    if (!h_batch) CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&h_batch), 8 * sizeof(bltk::SweepArgs), cudaHostAllocMapped));
    if (chunk_bytes > lane_chunk) release_lanes();
    lane_chunk = std::max(lane_chunk, chunk_bytes);
    while (lanes.size() < n_lanes) {
        lanes.emplace_back();
        GenLane &l = lanes.back();
        const size_t elems = std::max<size_t>(lane_chunk, 1u << 20);
        CUDA_TRY(cudaMalloc(&l.d_scratch, bltk::sweep_scratch_bytes(elems)));
        l.scratch = bltk::sweep_scratch_carve(l.d_scratch, elems);
        CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&l.d_work[0]), 2 * lane_chunk + 64));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&l.d_work[1]), 2 * lane_chunk + 64));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&l.d_align), lane_chunk + 64));
        CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&l.h_ctrl), 64, cudaHostAllocDefault));
    }
    return BLT_OK;
}
void Workspace::release() {
    release_lanes();
    if (d_scratch) cudaFree(d_scratch);
    for (auto &p : d_work) if (p) cudaFree(p);
    if (d_align) cudaFree(d_align);
    if (h_ctrl) cudaFreeHost(h_ctrl);
    if (h_batch) cudaFreeHost(h_batch);
    h_batch = nullptr;
    d_scratch = nullptr; d_work[0] = d_work[1] = nullptr; d_align = nullptr; h_ctrl = nullptr;
    scratch_elems = 0; work_chunk = 0;
}

// ---- run_device: enqueue the tokenization of a device buffer --------------------------------------
// Fixed-ratio strategies and the single-sweep byte-pair path are fully asynchronous; the general
// multi-sweep path synchronises `stream` once per sweep (it must read the "merged anything" flag).
int run_device(blt_strategy *s, Workspace &ws, const uint8_t *d_in, size_t n, size_t chunk, uint8_t *d_out,
               size_t out_cap, uint64_t *d_chunk_ends, cudaStream_t stream, DeviceResult *res) {
    res->kind = DeviceResult::KNOWN;
    res->len = 0;
    res->sweeps = 0;
    res->owner = nullptr;
    res->ratio_to = nullptr;
    res->n_in = 0;
    res->len_scale = 2;
    if (n == 0) return BLT_OK;
    if (chunk == 0 || chunk > n) chunk = n;
    if ((reinterpret_cast<uintptr_t>(d_in) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u))
        return fail(BLT_ERR_INVALID_INPUT, "device buffers must be 16-byte aligned");
    switch (s->mode) {
        case Mode::Basic: {
            if (out_cap < 2 * n) return fail(BLT_ERR_CAPACITY, "output capacity < 2*n");
            CUDA_TRY(bltk::launch_widen(d_in, n, d_out, stream));
            CUDA_TRY(bltk::launch_fill_chunk_ends(d_chunk_ends, n, chunk, 2, stream));
            res->len = 2 * n;
            res->launches = 1;
            return BLT_OK;
        }
        case Mode::Passthrough: {
            if (out_cap < n) return fail(BLT_ERR_CAPACITY, "output capacity < n");
            CUDA_TRY(cudaMemcpyAsync(d_out, d_in, n, cudaMemcpyDeviceToDevice, stream));
            CUDA_TRY(bltk::launch_fill_chunk_ends(d_chunk_ends, n, chunk, 1, stream));
            res->len = n;
            res->launches = 0;
            return BLT_OK;
        }
        case Mode::BpePairs: {
            int rc = ws.ensure_scratch(n);
            if (rc) return rc;
            bltk::SweepArgs a{};
            a.in = d_in; a.n = n; a.chunk = chunk;
            a.out = reinterpret_cast<uint16_t *>(d_out);
            a.out_cap_tokens = out_cap / 2; a.out_base_tokens = 0;
            a.chunk_ends = d_chunk_ends; a.chunk_ends_base = 0;
            a.scratch = ws.scratch;
            int variant = s->variant;
            if (!s->variant_forced) {
                const uint32_t r = s->last_ratio_milli.load(std::memory_order_relaxed);
                if (r != 0 && r < 515) variant = 0;
            }
            res->ratio_to = s;
            res->n_in = n;
            CUDA_TRY(bltk::launch_bpe_sweep_pairs(a, s->d_table, variant, s->want_dense(), stream, &res->launches));
            res->owner = (res->launches == bltk::kLaunchesDenseAttempt) ? s : nullptr;
            res->kind = DeviceResult::IN_SCRATCH;
            res->sweeps = 1;  // byte keys, ids >= 256: the reference's 2nd sweep cannot merge
            return BLT_OK;
        }
        case Mode::BpeGeneral: {
            // The loop of tokenizer.rs:63-86 (sweep until one merges nothing), for up to 8 chunks at a time: every chunk
            // of the batch launches its next sweep, ONE host synchronisation reads all their "merged anything" flags
            // and totals, chunks that are done drop out.  (Round 1 synchronised once per sweep and chunk.)
            const bltk::HashTableView view{s->d_slots, s->hash_mask, s->d_can_left, s->d_can_right, s->d_pair_bloom, s->d_bytemap};
            const size_t n_chunks = (n + chunk - 1) / chunk;
            size_t batch = std::min<size_t>(n_chunks, 8);
            while (batch > 1 && batch * 5 * chunk > (size_t(2) << 30)) --batch;  // at most 2 GiB of ping-pong buffers
            int rc = ws.ensure_lanes(chunk, batch);
            if (rc) return rc;
            std::vector<uint64_t> ends(n_chunks);
            size_t out_bytes = 0;
            uint32_t max_sweeps = 0;
            res->launches = 0;
            struct Live { const void *cur; size_t n; bool u16, active; int which; uint32_t sweeps; };
            for (size_t k0 = 0; k0 < n_chunks; k0 += batch) {
                const size_t nb = std::min(batch, n_chunks - k0);
                std::vector<Live> live(nb);
                for (size_t j = 0; j < nb; ++j) {
                    const size_t len = std::min(chunk, n - (k0 + j) * chunk);
                    const uint8_t *src = d_in + (k0 + j) * chunk;
                    if (reinterpret_cast<uintptr_t>(src) & 15u) {  // only when chunk is not a multiple of 16
                        CUDA_TRY(cudaMemcpyAsync(ws.lanes[j].d_align, src, len, cudaMemcpyDeviceToDevice, stream));
                        src = ws.lanes[j].d_align;
                    }
                    live[j] = Live{src, len, false, true, 0, 0};
                }
                for (bool any = true; any;) {
                    // one level: the next sweep of every chunk that is still merging, as ONE batch of three launches (the
                    // chunks of a level read the same kind of input: bytes in the first sweep, tokens afterwards)
                    bltk::SweepArgs *d_batch = nullptr;
                    CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_batch), ws.h_batch, 0));
                    int n_act = 0;
                    bool u16 = false;
                    for (size_t j = 0; j < nb; ++j) {
                        if (!live[j].active) continue;
                        Workspace::GenLane &ln = ws.lanes[j];
                        bltk::SweepArgs a{};
                        a.in = live[j].cur; a.n = live[j].n; a.chunk = 0;
                        a.out = reinterpret_cast<uint16_t *>(ln.d_work[live[j].which]);
                        a.out_cap_tokens = chunk; a.out_base_tokens = 0;
                        a.chunk_ends = nullptr; a.chunk_ends_base = 0;
                        a.scratch = ln.scratch;
                        a.skip_unmerged_emit = live[j].u16 ? 1u : 0u;
                        u16 = live[j].u16;
                        ws.h_batch[n_act++] = a;
                    }
                    CUDA_TRY(bltk::launch_bpe_sweep_hash_batch(ws.h_batch, d_batch, n_act, view, u16, stream));
                    for (size_t j = 0; j < nb; ++j) {
                        if (!live[j].active) continue;
                        Workspace::GenLane &ln = ws.lanes[j];
                        CUDA_TRY(cudaMemcpyAsync(ln.h_ctrl, ln.scratch.ctrl, 32, cudaMemcpyDeviceToHost, stream));
                        ++res->launches;
                    }
                    CUDA_TRY(cudaStreamSynchronize(stream));
                    any = false;
                    for (size_t j = 0; j < nb; ++j) {
                        if (!live[j].active) continue;
                        Workspace::GenLane &ln = ws.lanes[j];
                        ++live[j].sweeps;
                        const uint64_t total = ln.h_ctrl[0];
                        const uint32_t merged = reinterpret_cast<const uint32_t *>(ln.h_ctrl)[3];
                        if (reinterpret_cast<const uint32_t *>(ln.h_ctrl)[4]) return fail(BLT_ERR_CAPACITY, "output capacity exceeded");
                        // a sweep over tokens that merges nothing writes nothing (skip_unmerged_emit): its input is the
                        // result; the first sweep always writes, it widens the bytes
                        if (merged || !live[j].u16) {
                            live[j].cur = ln.d_work[live[j].which];
                            live[j].n = size_t(total);
                            live[j].u16 = true;
                            live[j].which ^= 1;
                        }
                        if (!merged) live[j].active = false;  // tokenizer.rs:83-85
                        else any = true;
                    }
                }
                for (size_t j = 0; j < nb; ++j) {
                    if (out_bytes + 2 * live[j].n > out_cap) return fail(BLT_ERR_CAPACITY, "output capacity exceeded");
                    CUDA_TRY(cudaMemcpyAsync(d_out + out_bytes, live[j].cur, 2 * live[j].n, cudaMemcpyDeviceToDevice, stream));
                    out_bytes += 2 * live[j].n;
                    ends[k0 + j] = out_bytes;
                    max_sweeps = std::max(max_sweeps, live[j].sweeps);
                }
            }
            if (d_chunk_ends) {
                CUDA_TRY(cudaMemcpyAsync(d_chunk_ends, ends.data(), n_chunks * 8, cudaMemcpyHostToDevice, stream));
                CUDA_TRY(cudaStreamSynchronize(stream));  // `ends` is a stack-owned host buffer
            }
            res->len = out_bytes;
            res->sweeps = max_sweeps;
            return BLT_OK;
        }
    }
    return fail(BLT_ERR_INVALID_INPUT, "unknown strategy mode");
}

// Reads back the result of a K2 launch (after the caller synchronised the stream that ran it, or
// by synchronising here).
int finish_result(Workspace &ws, cudaStream_t stream, DeviceResult *res) {
    if (res->kind != DeviceResult::IN_SCRATCH) return BLT_OK;
    if (!ws.h_ctrl) CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&ws.h_ctrl), 64, cudaHostAllocDefault));
    CUDA_TRY(cudaMemcpyAsync(ws.h_ctrl, ws.scratch.ctrl, 32, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return decode_ctrl(ws.h_ctrl, res);
}
int decode_ctrl(const uint64_t *h_ctrl, DeviceResult *res) {
    const uint32_t overflow = reinterpret_cast<const uint32_t *>(h_ctrl)[4];
    if (overflow == 2u) return fail(BLT_ERR_CUDA, "device-side launch of the exact sweep was refused");
    if (overflow == 3u) return fail(BLT_ERR_CUDA, "a single-pass kernel timed out waiting for a tile (look-back or bulk copy)");
    if (overflow) return fail(BLT_ERR_CAPACITY, "output capacity exceeded");
    if (res->owner) {  // the dense pass was attempted: tell the predictor how it went
        res->owner->dense_feedback(reinterpret_cast<const uint32_t *>(h_ctrl)[5] != 0u, reinterpret_cast<const uint32_t *>(h_ctrl)[7]);
        res->owner = nullptr;
    }
    if (res->len_scale == 1 && reinterpret_cast<const uint32_t *>(h_ctrl)[6] != 0u)
        return fail(BLT_ERR_INVALID_DATA, "token stream contains a token that is not in the table");
    res->len = size_t(h_ctrl[0]) * res->len_scale;
    if (res->ratio_to && res->n_in) {
        res->ratio_to->last_ratio_milli.store(uint32_t(std::max<uint64_t>(1, h_ctrl[0] * 1000 / res->n_in)), std::memory_order_relaxed);
        res->ratio_to = nullptr;
    }
    res->kind = DeviceResult::KNOWN;
    return BLT_OK;
}

// ---- detokenizer (no reference counterpart; see blt_cuda.h) ----------------------------------------
int run_detok(blt_strategy *s, Workspace &ws, const uint8_t *d_tokens, size_t n_bytes, uint8_t *d_out, size_t out_cap,
              cudaStream_t stream, DeviceResult *res) {
    res->kind = DeviceResult::KNOWN;
    res->len = 0;
    res->sweeps = 0;
    res->owner = nullptr;
    res->ratio_to = nullptr;
    res->n_in = 0;
    res->len_scale = 1;
    res->launches = 0;
    if (n_bytes & 1) return fail(BLT_ERR_INVALID_DATA, "token stream has an odd number of bytes");
    if (n_bytes == 0) return BLT_OK;
    if ((reinterpret_cast<uintptr_t>(d_tokens) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u))
        return fail(BLT_ERR_INVALID_INPUT, "device buffers must be 16-byte aligned");
    if (s->mode == Mode::Passthrough) {
        if (out_cap < n_bytes) return fail(BLT_ERR_CAPACITY, "output capacity exceeded");
        CUDA_TRY(cudaMemcpyAsync(d_out, d_tokens, n_bytes, cudaMemcpyDeviceToDevice, stream));
        res->len = n_bytes;
        return BLT_OK;
    }
    int rc = s->ensure_detok();
    if (rc) return rc;
    rc = ws.ensure_scratch(n_bytes / 2);  // the fused form keeps one 8-byte descriptor per 32 768 tokens in the scratch meta
    if (rc) return rc;
    bltk::DetokArgs a{};
    a.in = reinterpret_cast<const uint16_t *>(d_tokens);
    a.n_tok = n_bytes / 2;
    a.out = d_out;
    a.out_cap = out_cap;
    a.table = s->d_detok;
    a.limit = s->detok_limit;
    a.holes = s->detok_holes;
    a.scratch = ws.scratch;
    int launches = 0;
    CUDA_TRY(bltk::launch_detokenize(a, s->detok_variant, stream, &launches));
    res->kind = DeviceResult::IN_SCRATCH;
    res->launches = launches;
    return BLT_OK;
}

}  // namespace bltc

void blt_strategy::settle_pending() {
    // resident_mu is held.  Waits for the pending call's stream, reads its result back and keeps it for its thread.
    bltc::DeviceResult r = resident_result;
    r.rc = bltc::finish_result(resident, resident_stream, &r);
    if (r.rc == BLT_OK && cudaStreamSynchronize(resident_stream) != cudaSuccess) r.rc = bltc::fail(BLT_ERR_CUDA, "cudaStreamSynchronize failed");
    if (r.rc != BLT_OK) r.err = bltc::g_last_error;
    resident_settled[resident_owner] = r;
    resident_pending = false;
}

bool blt_strategy::want_dense() {
    if (!try_dense) return false;
    if (dense_always) return true;
    const uint32_t k = dense_skip.load(std::memory_order_relaxed);
    if (k != 0) { dense_skip.store(k - 1, std::memory_order_relaxed); return false; }
    // a probe: until it is known to have succeeded, assume it fails like the last one did
    dense_skip.store(dense_backoff.load(std::memory_order_relaxed), std::memory_order_relaxed);
    return true;
}
void blt_strategy::dense_feedback(bool failed, uint32_t prefix_permille) {
    // A failed attempt keeps the dense output of the chunks in front of the first failed one and hands only the rest to
    // the exact sweep, so an attempt that got through a good part of the launch was worth it: keep attempting.
    if (!failed || prefix_permille >= 250u) { dense_backoff.store(0, std::memory_order_relaxed); dense_skip.store(0, std::memory_order_relaxed); return; }
    const uint32_t b = dense_backoff.load(std::memory_order_relaxed);
    const uint32_t nb = b ? std::min(2 * b, 1024u) : 16u;
    dense_backoff.store(nb, std::memory_order_relaxed);
    dense_skip.store(nb, std::memory_order_relaxed);
}

int blt_strategy::ensure_detok() {
    std::lock_guard<std::mutex> lk(detok_mu);
    const char *not_inv = "this strategy's table is not invertible (keys must be byte pairs, ids >= 256 and distinct)";
    if (detok_state == 1) return BLT_OK;
    if (detok_state < 0) return bltc::fail(detok_state, not_inv);
    if (mode == bltc::Mode::BpeGeneral) {
        detok_state = BLT_ERR_INVALID_INPUT;
        return bltc::fail(detok_state, not_inv);
    }
    std::vector<uint16_t> dec(bltk::kPairTableEntries, 0);
    std::vector<uint32_t> exists(2048, 0);
    for (uint32_t b = 0; b < 256; ++b) dec[b] = uint16_t(b);  // a plain byte decodes to itself: one lookup whatever the width
    for (uint32_t w = 0; w < 8; ++w) exists[w] = 0xffffffffu;
    // later duplicates of a key overwrite (HashMap::insert, config_loader.rs:39): invert the FINAL map
    std::vector<int32_t> final_id(65536, -1);
    for (const auto &r : rules) final_id[size_t(r.left) | (size_t(r.right) << 8)] = int32_t(r.value);
    uint32_t max_id = 255, count = 0;
    for (uint32_t key = 0; key < 65536; ++key) {
        if (final_id[key] < 0) continue;
        const uint32_t id = uint32_t(final_id[key]);
        if ((exists[id >> 5] >> (id & 31)) & 1u) {
            detok_state = BLT_ERR_INVALID_INPUT;
            return bltc::fail(detok_state, not_inv);
        }
        exists[id >> 5] |= 1u << (id & 31);
        dec[id] = uint16_t(key);  // l | r << 8
        max_id = std::max(max_id, id);
        ++count;
    }
    detok_limit = max_id + 1;
    detok_holes = (count != max_id - 255) ? 1u : 0u;
    if (cudaSetDevice(ctx->device) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void **>(&d_detok), dec.size() * 2 + exists.size() * 4) != cudaSuccess)
        return bltc::fail(BLT_ERR_NOMEM, "cudaMalloc failed for the detokenizer table");
    if (cudaMemcpy(d_detok, dec.data(), dec.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(reinterpret_cast<unsigned char *>(d_detok) + dec.size() * 2, exists.data(), exists.size() * 4,
                   cudaMemcpyHostToDevice) != cudaSuccess)
        return bltc::fail(BLT_ERR_CUDA, "copying the detokenizer table failed");
    detok_state = 1;
    return BLT_OK;
}

namespace bltc {

// ---- strategy construction ------------------------------------------------------------------------
static int build_strategy(blt_ctx *ctx, blth::MergeList rules, blt_strategy **out) {
    auto s = std::unique_ptr<blt_strategy>(new blt_strategy());
    s->ctx = ctx;
    ctx->retain();
    s->rules = std::move(rules);
    CUDA_TRY(cudaSetDevice(ctx->device));
    bool pairs_ok = true;
    for (const auto &r : s->rules) pairs_ok = pairs_ok && r.left < 256 && r.right < 256 && r.value >= 256;
    if (pairs_ok) {
        // Every key is a byte pair and every id is >= 256 (always true for a merges.txt,
        // config_loader.rs:27-40): one sweep is the fixpoint and the table is direct-indexed.
        s->mode = Mode::BpePairs;
        std::vector<uint16_t> tbl(bltk::kPairTableEntries);
        for (uint32_t b0 = 0; b0 < 256; ++b0)
            for (uint32_t b1 = 0; b1 < 256; ++b1) tbl[bltk::pair_table_index(b0, b1)] = uint16_t(b0 << 8);
        for (const auto &r : s->rules)
            tbl[bltk::pair_table_index(r.left, r.right)] = uint16_t((r.value >> 8) | (r.value << 8));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&s->d_table), tbl.size() * 2));
        CUDA_TRY(cudaMemcpy(s->d_table, tbl.data(), tbl.size() * 2, cudaMemcpyHostToDevice));
    } else {
        s->mode = Mode::BpeGeneral;
        size_t cap = 16;
        while (cap < 2 * s->rules.size() + 2) cap <<= 1;
        std::vector<bltk::HashSlot> slots(cap, bltk::HashSlot{0, 0, 0});
        std::vector<uint32_t> cl(2048, 0), cr(2048, 0), bloom(2048, 0);
        for (const auto &r : s->rules) {
            const uint32_t key = (uint32_t(r.left) << 16) | r.right;
            uint32_t h = bltk::hash_pair(key) & uint32_t(cap - 1);
            while (slots[h].used) h = (h + 1) & uint32_t(cap - 1);
            slots[h] = bltk::HashSlot{key, r.value, 1};
            cl[r.left >> 5] |= 1u << (r.left & 31);
            cr[r.right >> 5] |= 1u << (r.right & 31);
            const uint32_t bb = bltk::pair_bloom_bit(r.left, r.right);
            bloom[bb >> 5] |= 1u << (bb & 31);
        }
        s->hash_mask = uint32_t(cap - 1);
        CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&s->d_slots), cap * sizeof(bltk::HashSlot)));
        CUDA_TRY(cudaMemcpy(s->d_slots, slots.data(), cap * sizeof(bltk::HashSlot), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&s->d_can_left), 8192));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&s->d_can_right), 8192));
        CUDA_TRY(cudaMemcpy(s->d_can_left, cl.data(), 8192, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(s->d_can_right, cr.data(), 8192, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&s->d_pair_bloom), 8192));
        CUDA_TRY(cudaMemcpy(s->d_pair_bloom, bloom.data(), 8192, cudaMemcpyHostToDevice));
        // the first sweep reads bytes: the rules with byte components as a direct table (values + exact membership bitmap)
        std::vector<uint16_t> bytemap(bltk::kPairTableEntries + 4096, 0);
        uint32_t *member = reinterpret_cast<uint32_t *>(bytemap.data() + bltk::kPairTableEntries);
        for (const auto &r : s->rules) {
            if (r.left > 255 || r.right > 255) continue;
            const uint32_t idx = bltk::pair_table_index(r.left, r.right);
            bytemap[idx] = r.value;  // (the list holds the final map: one rule per key)
            member[idx >> 5] |= 1u << (idx & 31);
        }
        CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&s->d_bytemap), bytemap.size() * 2));
        CUDA_TRY(cudaMemcpy(s->d_bytemap, bytemap.data(), bytemap.size() * 2, cudaMemcpyHostToDevice));
    }
    // tuning / test switches: tile size of the exact sweep, and whether the dense pass runs in front of it
    if (const char *v = getenv("BLT_SWEEP_VARIANT")) { s->variant = atoi(v); s->variant_forced = true; }
    if (const char *v = getenv("BLT_DETOK_VARIANT")) s->detok_variant = atoi(v);
    if (const char *v = getenv("BLT_DENSE")) {  // 0 = never, always = on every call, else the predictor decides
        s->dense_always = std::string(v) == "always";
        s->try_dense = s->dense_always || atoi(v) != 0;
    }
    *out = s.release();
    return BLT_OK;
}

}  // namespace bltc

blt_strategy::~blt_strategy() {
    if (ctx) cudaSetDevice(ctx->device);
    if (d_table) cudaFree(d_table);
    if (d_slots) cudaFree(d_slots);
    if (d_can_left) cudaFree(d_can_left);
    if (d_can_right) cudaFree(d_can_right);
    if (d_pair_bloom) cudaFree(d_pair_bloom);
    if (d_bytemap) cudaFree(d_bytemap);
    if (d_detok) cudaFree(d_detok);
    resident.release();
    if (ctx) ctx->release_ref();
}

#ifdef BLT_FUSED_PROF
namespace bltk { cudaError_t debug_fused_profile(unsigned long long *host_out, size_t n_words); }
extern "C" __attribute__((visibility("default"))) int blt_debug_fused_profile(unsigned long long *out, size_t n_words) {
    return bltk::debug_fused_profile(out, n_words) == cudaSuccess ? 0 : -5;
}
#endif

using namespace bltc;

extern "C" {

const char *blt_version(void) { return BLT_VERSION_STRING; }
const char *blt_last_error(void) { return g_last_error.c_str(); }

int blt_device_count(int *count) {
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        if (count) *count = 0;
        cudaGetLastError();
        return fail(BLT_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
    }
    if (count) *count = n;
    return BLT_OK;
}

int blt_ctx_create(int device, blt_ctx **out) {
    if (!out) return fail(BLT_ERR_INVALID_INPUT, "out is NULL");
    *out = nullptr;
    int n = 0;
    int rc = blt_device_count(&n);
    if (rc) return rc;
    if (device < 0 || device >= n) return fail(BLT_ERR_INVALID_INPUT, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaFree(nullptr));  // force context creation so errors surface here
    auto c = new blt_ctx();
    c->device = device;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = c;
    return BLT_OK;
}

void blt_ctx_destroy(blt_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        for (auto &p : ctx->idle) p->release();
        ctx->idle.clear();
    }
    ctx->release_ref();
}

int blt_strategy_basic(blt_ctx *ctx, blt_strategy **out) {
    if (!ctx || !out) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    auto s = new blt_strategy();
    s->ctx = ctx;
    ctx->retain();
    s->mode = Mode::Basic;
    *out = s;
    return BLT_OK;
}

int blt_strategy_passthrough(blt_ctx *ctx, blt_strategy **out) {
    if (!ctx || !out) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    auto s = new blt_strategy();
    s->ctx = ctx;
    ctx->retain();
    s->mode = Mode::Passthrough;
    *out = s;
    return BLT_OK;
}

int blt_strategy_bpe_from_file(blt_ctx *ctx, const char *merges_path, blt_strategy **out) {
    if (!ctx || !out || !merges_path) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    *out = nullptr;
    blth::MergeList rules;
    const blth::Error e = blth::load_merges_file(merges_path, &rules);
    if (e) return fail(BLT_ERR_INVALID_INPUT, "Failed to load BPE merges: " + e.msg);  // lib.rs:194-201
    return build_strategy(ctx, std::move(rules), out);
}

int blt_strategy_bpe_from_pairs(blt_ctx *ctx, const uint16_t *left, const uint16_t *right, const uint16_t *value,
                                size_t n, blt_strategy **out) {
    if (!ctx || !out || (n && (!left || !right || !value))) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    *out = nullptr;
    std::vector<blth::MergeRule> rules(n);
    for (size_t i = 0; i < n; ++i) rules[i] = blth::MergeRule{left[i], right[i], value[i]};
    return build_strategy(ctx, blth::dedup_rules(rules), out);
}

void blt_strategy_destroy(blt_strategy *s) { delete s; }

size_t blt_strategy_num_merges(const blt_strategy *s) { return s ? s->rules.size() : 0; }

int blt_process_chunk(blt_strategy *s, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len) {
    if (!s || !out_len || (n && (!in || !out))) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    *out_len = 0;
    if (n == 0) return BLT_OK;  // tokenizer.rs:57-59, 109-111
    return bltc::tokenize_host(s, in, n, n, BLT_CONTENT_NONE, out, out_cap, out_len);
}

int blt_tokenize_host(blt_strategy *s, const uint8_t *in, size_t n, size_t chunk_size, int content_type, uint8_t *out,
                      size_t out_cap, size_t *out_len) {
    if (!s || !out_len || (n && !in) || !out) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    *out_len = 0;
    return bltc::tokenize_host(s, in, n, chunk_size, content_type, out, out_cap, out_len);
}

int blt_process_resident(blt_strategy *s, const void *d_in, size_t n, size_t chunk_size, void *d_out, size_t out_cap,
                         uint64_t *d_chunk_ends, void *stream, size_t *out_len) {
    if (!s || (n && (!d_in || !d_out))) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    std::lock_guard<std::mutex> lk(s->resident_mu);
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (s->resident_pending && (s->resident_owner != std::this_thread::get_id() || s->resident_stream != st)) s->settle_pending();
    s->resident_pending = false;
    int rc = run_device(s, s->resident, static_cast<const uint8_t *>(d_in), n, chunk_size, static_cast<uint8_t *>(d_out),
                        out_cap, d_chunk_ends, st, &s->resident_result);
    if (rc) return rc;
    if (out_len) {
        rc = finish_result(s->resident, st, &s->resident_result);
        if (rc) return rc;
        if (s->resident_result.kind == DeviceResult::KNOWN) CUDA_TRY(cudaStreamSynchronize(st));
        *out_len = s->resident_result.len;
        return BLT_OK;
    }
    s->resident_pending = true;
    s->resident_owner = std::this_thread::get_id();
    s->resident_stream = st;
    s->resident_settled.erase(s->resident_owner);
    return BLT_OK;
}

int blt_detokenize_host(blt_strategy *s, const uint8_t *in, size_t n_bytes, int has_content_type, uint8_t *out,
                        size_t out_cap, size_t *out_len) {
    if (!s || !out_len || (n_bytes && !in) || !out) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    *out_len = 0;
    if (n_bytes & 1) return fail(BLT_ERR_INVALID_DATA, "token stream has an odd number of bytes");
    if (has_content_type) {  // prepend_content_type_token's inverse (lib.rs:284-294)
        if (n_bytes < 2) return fail(BLT_ERR_INVALID_DATA, "missing content-type token");
        const uint32_t t = (uint32_t(in[0]) << 8) | in[1];
        if (t < 0xFF01u || t > 0xFF04u) return fail(BLT_ERR_INVALID_DATA, "missing content-type token");
        in += 2;
        n_bytes -= 2;
    }
    return bltc::detokenize_host(s, in, n_bytes, out, out_cap, out_len);
}

int blt_detokenize_resident(blt_strategy *s, const void *d_tokens, size_t n_bytes, void *d_out, size_t out_cap,
                            void *stream, size_t *out_len) {
    if (!s || (n_bytes && (!d_tokens || !d_out))) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    std::lock_guard<std::mutex> lk(s->resident_mu);
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (s->resident_pending && (s->resident_owner != std::this_thread::get_id() || s->resident_stream != st)) s->settle_pending();
    s->resident_pending = false;
    int rc = run_detok(s, s->resident, static_cast<const uint8_t *>(d_tokens), n_bytes, static_cast<uint8_t *>(d_out), out_cap,
                       st, &s->resident_result);
    if (rc) return rc;
    if (out_len) {
        rc = finish_result(s->resident, st, &s->resident_result);
        if (rc) return rc;
        if (s->resident_result.kind == DeviceResult::KNOWN) CUDA_TRY(cudaStreamSynchronize(st));
        *out_len = s->resident_result.len;
        return BLT_OK;
    }
    s->resident_pending = true;
    s->resident_owner = std::this_thread::get_id();
    s->resident_stream = st;
    s->resident_settled.erase(s->resident_owner);
    return BLT_OK;
}

int blt_count_pairs_resident(blt_ctx *ctx, const void *d_in, size_t n, uint64_t *d_counts, void *stream) {
    if (!ctx || !d_counts || (n && !d_in)) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    if (reinterpret_cast<uintptr_t>(d_in) & 15u) return fail(BLT_ERR_INVALID_INPUT, "device buffers must be 16-byte aligned");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(bltk::launch_pair_hist(static_cast<const unsigned char *>(d_in), n, reinterpret_cast<unsigned long long *>(d_counts),
                                    true, static_cast<cudaStream_t>(stream)));
    return BLT_OK;
}

int blt_count_pairs_host(blt_ctx *ctx, const uint8_t *in, size_t n, uint64_t *counts) {
    if (!ctx || !counts || (n && !in)) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t unit = size_t(256) << 20;  // pairs START inside a unit; one byte of the next unit rides along
    unsigned char *d_buf = nullptr;
    unsigned long long *d_counts = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&d_counts), 65536 * sizeof(unsigned long long)));
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&d_buf), std::min(n, unit) + 16);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_counts, 0, 65536 * sizeof(unsigned long long), nullptr);
    for (size_t off = 0; e == cudaSuccess && off < n; off += unit) {
        const size_t len = std::min(unit + 1, n - off);
        e = cudaMemcpy(d_buf, in + off, len, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = bltk::launch_pair_hist(d_buf, len, d_counts, false, nullptr);
        if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    }
    if (e == cudaSuccess) e = cudaMemcpy(counts, d_counts, 65536 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (d_buf) cudaFree(d_buf);
    cudaFree(d_counts);
    if (e != cudaSuccess) return fail(BLT_ERR_CUDA, std::string("pair histogram failed: ") + cudaGetErrorString(e));
    return BLT_OK;
}

int blt_select_merges(const uint64_t *counts, size_t k, int pad_unobserved, uint8_t *left, uint8_t *right, size_t *n_out) {
    if (!counts || !n_out || (k && (!left || !right))) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    if (k > 65280) return fail(BLT_ERR_INVALID_INPUT, "at most 65280 rules fit the u16 id space (config_loader.rs:18,40)");
    std::vector<uint32_t> keys;
    for (uint32_t key = 0; key < 65536; ++key)
        if (counts[key]) keys.push_back(key);
    std::sort(keys.begin(), keys.end(), [&](uint32_t a, uint32_t b) { return counts[a] != counts[b] ? counts[a] > counts[b] : a < b; });
    size_t w = 0;
    for (size_t i = 0; i < keys.size() && w < k; ++i, ++w) { left[w] = uint8_t(keys[i] >> 8); right[w] = uint8_t(keys[i] & 0xff); }
    if (pad_unobserved)
        for (uint32_t key = 0; key < 65536 && w < k; ++key)
            if (!counts[key]) { left[w] = uint8_t(key >> 8); right[w] = uint8_t(key & 0xff); ++w; }
    *n_out = w;
    return BLT_OK;
}

int blt_resident_result(blt_strategy *s, void *stream, size_t *out_len, uint32_t *sweeps) {
    if (!s) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    std::lock_guard<std::mutex> lk(s->resident_mu);
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const std::thread::id me = std::this_thread::get_id();
    if (!(s->resident_pending && s->resident_owner == me)) {
        // settled on this thread's behalf by a later call of another thread, or collected before
        auto it = s->resident_settled.find(me);
        if (it == s->resident_settled.end()) {
            if (s->resident_pending) return fail(BLT_ERR_INVALID_INPUT, "no device-resident call of this thread is outstanding");
            // no asynchronous call outstanding: the most recent (synchronous or collected) call of the strategy
            if (out_len) *out_len = s->resident_result.len;
            if (sweeps) *sweeps = s->resident_result.sweeps;
            return BLT_OK;
        }
        const DeviceResult r = it->second;
        if (r.rc != BLT_OK) return fail(r.rc, r.err);
        if (out_len) *out_len = r.len;
        if (sweeps) *sweeps = r.sweeps;
        return BLT_OK;
    }
    int rc = finish_result(s->resident, st, &s->resident_result);
    if (rc) { s->resident_pending = false; return rc; }
    CUDA_TRY(cudaStreamSynchronize(st));
    s->resident_pending = false;
    if (out_len) *out_len = s->resident_result.len;
    if (sweeps) *sweeps = s->resident_result.sweeps;
    return BLT_OK;
}

// ---- host-only helpers ----------------------------------------------------------------------------
int blt_load_bpe_merges(const char *path, uint16_t *left, uint16_t *right, uint16_t *value, size_t cap, size_t *n) {
    if (!path || !n) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    blth::MergeList rules;
    const blth::Error e = blth::load_merges_file(path, &rules);
    if (e) return fail(e.code, e.msg);
    *n = rules.size();
    for (size_t i = 0; i < rules.size() && i < cap; ++i) {
        if (left) left[i] = rules[i].left;
        if (right) right[i] = rules[i].right;
        if (value) value[i] = rules[i].value;
    }
    return BLT_OK;
}

int blt_parse_chunk_size(const char *s, size_t *out) {
    if (!s || !out) return fail(BLT_ERR_INVALID_INPUT, "NULL argument");
    const blth::Error e = blth::parse_chunk_size(s, out);
    if (e) return fail(e.code, e.msg);
    return BLT_OK;
}

size_t blt_effective_chunk_size(int has_cli, size_t cli_size, size_t threads, unsigned memcap, uint64_t total_ram_bytes) {
    return blth::effective_chunk_size(has_cli != 0, cli_size, threads, memcap, total_ram_bytes);
}

size_t blt_determine_thread_count(int has_override, size_t override_val) {
    return blth::determine_thread_count(has_override != 0, override_val);
}

uint16_t blt_content_type_token(int ct) {
    switch (ct) {  // lib.rs:96-103
        case BLT_CONTENT_TEXT: return 0xFF01;
        case BLT_CONTENT_AUDIO: return 0xFF02;
        case BLT_CONTENT_BIN: return 0xFF03;
        case BLT_CONTENT_VIDEO: return 0xFF04;
        default: return 0;
    }
}

void blt_shard_chunks(size_t n_chunks, int n_gpus, size_t *bounds) {
    if (n_gpus < 1) n_gpus = 1;
    // chunk k -> GPU floor(k*G/K): GPU g owns [ceil(g*K/G), ceil((g+1)*K/G))
    for (int g = 0; g <= n_gpus; ++g) bounds[g] = (size_t(g) * n_chunks + size_t(n_gpus) - 1) / size_t(n_gpus);
}

int blt_file_chunk_device(size_t chunk_index, int n_gpus) { return n_gpus > 1 ? int(chunk_index % size_t(n_gpus)) : 0; }

}  // extern "C"
