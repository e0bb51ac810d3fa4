// pipeline.h -- internal types shared by cabi.cu and pipeline.cu (not part of the C ABI).
#pragma once

#include "../../include/blt_cuda.h"
#include "host_config.h"
#include "kernels.cuh"

#include <atomic>
#include <map>
#include <memory>
#include <thread>
#include <mutex>
#include <string>
#include <vector>

namespace bltc {

enum class Mode { Basic, Passthrough, BpePairs, BpeGeneral };

int fail(int code, const std::string &msg);

// Device scratch for one in-flight tokenization (control block, range descriptors, multi-sweep ping-pong).
struct Workspace {
    void *d_scratch = nullptr;
    size_t scratch_elems = 0;
    bltk::SweepScratch scratch{};
    uint8_t *d_work[2] = {nullptr, nullptr};  // general path: sweep ping-pong, 2*chunk bytes each
    uint8_t *d_align = nullptr;               // general path: 16-byte aligned copy of a chunk
    size_t work_chunk = 0;
    uint64_t *h_ctrl = nullptr;               // pinned mirror of the scratch control block
    // general (multi-sweep) path: several chunks advance together, one host synchronisation per sweep LEVEL of the batch
    // instead of one per sweep and chunk; every chunk of the batch needs its own scratch, ping-pong buffers and mirror
    struct GenLane {
        void *d_scratch = nullptr;
        bltk::SweepScratch scratch{};
        uint8_t *d_work[2] = {nullptr, nullptr};
        uint8_t *d_align = nullptr;
        uint64_t *h_ctrl = nullptr;
    };
    std::vector<GenLane> lanes;
    size_t lane_chunk = 0;
    bltk::SweepArgs *h_batch = nullptr;       // page-locked, device-readable: the arguments of the sweeps of one level (<= 8)
    int ensure_scratch(size_t n_elems);
    int ensure_work(size_t chunk_bytes);
    int ensure_lanes(size_t chunk_bytes, size_t n_lanes);
    void release_lanes();
    void release();
};

struct DeviceResult {
    enum Kind { KNOWN, IN_SCRATCH } kind = KNOWN;  // IN_SCRATCH: length still in device memory
    size_t len = 0;       // output bytes
    uint32_t sweeps = 0;  // productive + verifying sweeps actually launched
    int launches = 0;     // kernels enqueued by the host
    blt_strategy *owner = nullptr;  // set when the dense pass was attempted: decode_ctrl reports back to it
    blt_strategy *ratio_to = nullptr;  // byte-pair sweeps: decode_ctrl files tokens-per-byte of this call there
    size_t n_in = 0;                   //   (input elements of the call)
    uint32_t len_scale = 2;         // bytes per unit of the device's total (2: tokens, 1: detokenizer bytes)
    int rc = 0;                     // a result settled on its caller's behalf: the code and message it ended with
    std::string err;
};

int run_device(blt_strategy *s, Workspace &ws, const uint8_t *d_in, size_t n, size_t chunk, uint8_t *d_out,
               size_t out_cap, uint64_t *d_chunk_ends, cudaStream_t stream, DeviceResult *res);
int run_detok(blt_strategy *s, Workspace &ws, const uint8_t *d_tokens, size_t n_bytes, uint8_t *d_out, size_t out_cap,
              cudaStream_t stream, DeviceResult *res);
int finish_result(Workspace &ws, cudaStream_t stream, DeviceResult *res);
int decode_ctrl(const uint64_t *h_ctrl, DeviceResult *res);

// One slot of the chunk pipeline: device in/out buffers plus the pinned words the kernel result is
// copied into.  Pinned staging (h_in/h_out) is only allocated by the file-to-file pipeline.
struct Slot {
    uint8_t *d_in = nullptr, *d_out = nullptr;
    uint8_t *h_in = nullptr, *h_out = nullptr;
    uint64_t *h_ctrl = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_done = nullptr, ev_d2h = nullptr;
    size_t in_len = 0;
    DeviceResult res;
};

// The resources of one concurrent host call: three streams (H2D, compute, D2H) and S slots.
struct Pipe {
    int device = 0;
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
    std::vector<Slot> slots;
    size_t cap = 0;        // input bytes per slot (device buffers)
    size_t pin_cap = 0;    // input bytes per slot of the pinned staging (0: none)
    std::vector<cudaEvent_t> piece_ev;  // single-unit calls: one event per output piece (see run_one_unit_pieced)
    Workspace ws;
    int ensure(size_t chunk_cap, size_t n_slots, bool want_pinned);  // want_pinned: staging for chunk_cap bytes too
    void release();
};

int tokenize_host(blt_strategy *s, const uint8_t *in, size_t n, size_t chunk, int content_type, uint8_t *out,
                  size_t out_cap, size_t *out_len);
int detokenize_host(blt_strategy *s, const uint8_t *in, size_t n_bytes, uint8_t *out, size_t out_cap, size_t *out_len);

}  // namespace bltc

struct blt_ctx {
    int device = 0;
    int sm_count = 0;
    // 1 for the handle the caller holds + 1 per live strategy: blt_ctx_destroy with strategies still alive only gives
    // the pooled pipes back; the object goes away with the last strategy (a strategy dereferences its context)
    std::atomic<int> refs{1};
    void retain() { refs.fetch_add(1, std::memory_order_relaxed); }
    void release_ref() { if (refs.fetch_sub(1, std::memory_order_acq_rel) == 1) delete this; }
    std::mutex mu;
    std::vector<std::unique_ptr<bltc::Pipe>> idle;  // pipes not currently lent to a call
    std::unique_ptr<bltc::Pipe> acquire();
    void give_back(std::unique_ptr<bltc::Pipe> p);
};

struct blt_strategy {
    blt_ctx *ctx = nullptr;
    bltc::Mode mode = bltc::Mode::Basic;
    blth::MergeList rules;
    uint16_t *d_table = nullptr;          // K2 byte-pair table
    bltk::HashSlot *d_slots = nullptr;    // K3 hash table
    uint32_t *d_can_left = nullptr, *d_can_right = nullptr, *d_pair_bloom = nullptr;
    uint16_t *d_bytemap = nullptr;        // K3, first sweep: direct table of the rules with byte components
    uint32_t hash_mask = 0;
    int variant = 3;                      // K2 exact sweep: 3 = fused single pass (default), 0/1/2 = count/scan/emit forms, 4 = fused 23x2
    int detok_variant = 0;                // detokenizer: 0 = count/scan/emit (default, measured faster), 1 = fused single pass (BLT_DETOK_VARIANT)
    bool try_dense = true;                // K2: run the speculative dense pass first
    // Predictor of the speculation.  A failed attempt costs the attempt plus a device-side launch of the
    // exact sweep (~0.1 ms), so after a failure the next `backoff` calls go to the exact sweep directly
    // (16, doubling to 1024 while the probes keep failing); one success makes every call dense again.
    std::atomic<uint32_t> dense_skip{0}, dense_backoff{0};
    bool dense_always = false;            // BLT_DENSE=always: attempt on every call (tests)
    // Which exact sweep a byte-pair call takes when BLT_SWEEP_VARIANT does not say: the fused single pass, except that
    // input on which (nearly) every pair merges (the last call emitted < 0.515 tokens per byte) goes to the three-launch
    // form, which is faster there (0.91 against 1.00 ms per GiB: 32 warps per SM, no look-back) and slower elsewhere.
    bool variant_forced = false;
    std::atomic<uint32_t> last_ratio_milli{0};  // 1000 * tokens / input bytes of the most recent call, 0: none yet
    bool want_dense();
    void dense_feedback(bool failed, uint32_t prefix_permille = 0);
    std::mutex detok_mu;                  // detokenizer table, built on first use
    uint16_t *d_detok = nullptr;          //   65536 x u16 + 2048 x u32 (DetokArgs::table)
    uint32_t detok_limit = 0, detok_holes = 0;
    int detok_state = 0;                  //   0 not built, 1 ready, < 0 the error code it failed with
    int ensure_detok();
    // blt_process_resident / blt_detokenize_resident share ONE workspace (control block, descriptors).  An
    // asynchronous call (out_len == NULL) stays "pending" until its result is collected; a call from another thread
    // or on another stream first settles the pending one (waits for it, reads its result back and files it under
    // the thread that made it), so two calls never run on the scratch at once and blt_resident_result returns the
    // calling thread's own most recent call.  Calls of one thread on one stream are ordered by the stream itself.
    std::mutex resident_mu;
    bltc::Workspace resident;             // workspace of blt_process_resident
    bltc::DeviceResult resident_result;   // of the pending call
    bool resident_pending = false;
    std::thread::id resident_owner;
    cudaStream_t resident_stream = nullptr;
    std::map<std::thread::id, bltc::DeviceResult> resident_settled;
    void settle_pending();                // resident_mu held
    ~blt_strategy();
};
