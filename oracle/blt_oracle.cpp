// blt_oracle.cpp -- CPU ORACLE (test infrastructure, never shipped, never on the product path).
// See blt_oracle.h for the contract and the parity-pinning statement.
//
// Style note: the data-structure classes deliberately match the reference so that the cost
// profile is comparable when this file is timed as the CPU baseline ("port"): a hash map keyed by
// (u16,u16) hashed with SipHash-1-3 (what Rust's std HashMap does), a fresh token vector per
// sweep, a fresh output vector per chunk, <= T chunks in flight and one ordered writer.

#include "blt_oracle.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <mutex>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <vector>

namespace {

// ---- SipHash-1-3 (the hasher behind Rust's RandomState; keys fixed here) ------------------------
inline uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }

struct SipPairHash {
    // Hashes the 4 key bytes (left LE, right LE), as `(u16,u16)::hash` feeds them to the hasher.
    size_t operator()(uint32_t key) const noexcept {
        const uint64_t k0 = 0x0706050403020100ULL, k1 = 0x0f0e0d0c0b0a0908ULL;
        uint64_t v0 = k0 ^ 0x736f6d6570736575ULL, v1 = k1 ^ 0x646f72616e646f6dULL;
        uint64_t v2 = k0 ^ 0x6c7967656e657261ULL, v3 = k1 ^ 0x7465646279746573ULL;
        auto round = [&]() {
            v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32);
            v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;
            v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;
            v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32);
        };
        uint64_t b = (uint64_t(4) << 56) | uint64_t(key);  // length byte + the 4 message bytes
        v3 ^= b; round(); v0 ^= b;                          // 1 compression round
        v2 ^= 0xff; round(); round(); round();              // 3 finalisation rounds
        return size_t(v0 ^ v1 ^ v2 ^ v3);
    }
};

inline uint32_t pack(uint16_t l, uint16_t r) { return (uint32_t(l) << 16) | r; }

void set_err(char *err, size_t cap, const std::string &msg) {
    if (err && cap) {
        std::snprintf(err, cap, "%s", msg.c_str());
    }
}

}  // namespace

struct ora_merges {
    std::unordered_map<uint32_t, uint16_t, SipPairHash> map;
};

extern "C" {

ora_merges *ora_merges_new(void) { return new ora_merges(); }
void ora_merges_free(ora_merges *m) { delete m; }
void ora_merges_insert(ora_merges *m, uint16_t l, uint16_t r, uint16_t v) {
    m->map[pack(l, r)] = v;
}
size_t ora_merges_len(const ora_merges *m) { return m->map.size(); }

size_t ora_merges_export(const ora_merges *m, uint16_t *left, uint16_t *right, uint16_t *value,
                         size_t cap) {
    std::vector<std::pair<uint32_t, uint16_t>> v(m->map.begin(), m->map.end());
    std::sort(v.begin(), v.end());
    size_t n = std::min(cap, v.size());
    for (size_t i = 0; i < n; ++i) {
        left[i] = uint16_t(v[i].first >> 16);
        right[i] = uint16_t(v[i].first & 0xffff);
        value[i] = v[i].second;
    }
    return n;
}

}  // extern "C"

namespace {

// ---- merges.txt parsing helpers (config_loader.rs:14-46) ----------------------------------------

// Rust's `str::parse::<u8>()`: optional single leading '+', then >=1 ASCII digits, value <= 255.
// Returns "" on success, else the ParseIntError Display text.
std::string parse_u8(const std::string &s, uint8_t *out) {
    if (s.empty()) return "cannot parse integer from empty string";
    size_t i = 0;
    if (s[0] == '+' || s[0] == '-') {
        // A lone sign is an invalid digit; '-' is never accepted for unsigned targets.
        if (s.size() == 1 || s[0] == '-') return "invalid digit found in string";
        i = 1;
    }
    unsigned v = 0;
    for (; i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9') return "invalid digit found in string";
        v = v * 10 + unsigned(s[i] - '0');
        // Rust's checked loop reports the overflow at the digit where it happens, before it
        // looks at any later character.
        if (v > 255) return "number too large to fit in target type";
    }
    *out = uint8_t(v);
    return "";
}

// Decode one UTF-8 scalar; returns its length or 0 if invalid (matches Rust's from_utf8 rules:
// no overlongs, no surrogates, max U+10FFFF).
size_t utf8_decode(const unsigned char *p, size_t n, uint32_t *cp) {
    if (n == 0) return 0;
    unsigned char c = p[0];
    if (c < 0x80) { *cp = c; return 1; }
    if (c >= 0xC2 && c <= 0xDF) {
        if (n < 2 || (p[1] & 0xC0) != 0x80) return 0;
        *cp = (uint32_t(c & 0x1F) << 6) | (p[1] & 0x3F);
        return 2;
    }
    if (c >= 0xE0 && c <= 0xEF) {
        if (n < 3 || (p[1] & 0xC0) != 0x80 || (p[2] & 0xC0) != 0x80) return 0;
        if (c == 0xE0 && p[1] < 0xA0) return 0;
        if (c == 0xED && p[1] > 0x9F) return 0;
        *cp = (uint32_t(c & 0x0F) << 12) | (uint32_t(p[1] & 0x3F) << 6) | (p[2] & 0x3F);
        return 3;
    }
    if (c >= 0xF0 && c <= 0xF4) {
        if (n < 4 || (p[1] & 0xC0) != 0x80 || (p[2] & 0xC0) != 0x80 || (p[3] & 0xC0) != 0x80)
            return 0;
        if (c == 0xF0 && p[1] < 0x90) return 0;
        if (c == 0xF4 && p[1] > 0x8F) return 0;
        *cp = (uint32_t(c & 0x07) << 18) | (uint32_t(p[1] & 0x3F) << 12) |
              (uint32_t(p[2] & 0x3F) << 6) | (p[3] & 0x3F);
        return 4;
    }
    return 0;
}

// Unicode White_Space, the predicate behind `str::split_whitespace`.
bool is_unicode_ws(uint32_t c) {
    return (c >= 0x09 && c <= 0x0D) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 ||
           (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 || c == 0x202F ||
           c == 0x205F || c == 0x3000;
}

// Returns false if the line is not valid UTF-8.
bool split_whitespace(const std::string &line, std::vector<std::string> *parts) {
    parts->clear();
    const unsigned char *p = reinterpret_cast<const unsigned char *>(line.data());
    size_t n = line.size(), i = 0, tok_start = 0;
    bool in_tok = false;
    while (i < n) {
        uint32_t cp;
        size_t len = utf8_decode(p + i, n - i, &cp);
        if (len == 0) return false;
        if (is_unicode_ws(cp)) {
            if (in_tok) { parts->push_back(line.substr(tok_start, i - tok_start)); in_tok = false; }
        } else if (!in_tok) {
            in_tok = true;
            tok_start = i;
        }
        i += len;
    }
    if (in_tok) parts->push_back(line.substr(tok_start));
    return true;
}

int errno_to_kind(int e) { return e == ENOENT ? ORA_NOT_FOUND : ORA_IO; }

bool read_whole_file(const char *path, std::string *data, int *kind, std::string *msg) {
    FILE *f = std::fopen(path, "rb");
    if (!f) {
        *kind = errno_to_kind(errno);
        *msg = std::string(std::strerror(errno)) + " (os error " + std::to_string(errno) + ")";
        return false;
    }
    char buf[1 << 16];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) data->append(buf, got);
    std::fclose(f);
    return true;
}

}  // namespace

extern "C" int ora_load_bpe_merges(const char *path, ora_merges **out, char *err, size_t errcap) {
    *out = nullptr;
    std::string data, msg;
    int kind = ORA_OK;
    if (!read_whole_file(path, &data, &kind, &msg)) {  // File::open(path)?  config_loader.rs:15
        set_err(err, errcap, msg);
        return kind;
    }
    ora_merges *m = new ora_merges();
    uint32_t vocab_size = 256;  // config_loader.rs:18 (u16 in the reference, see overflow note below)
    size_t pos = 0;
    std::vector<std::string> parts;
    while (pos < data.size()) {  // BufRead::lines(): split on '\n', strip one trailing "\r"
        size_t nl = data.find('\n', pos);
        size_t end = (nl == std::string::npos) ? data.size() : nl;
        std::string line = data.substr(pos, end - pos);
        pos = (nl == std::string::npos) ? data.size() : nl + 1;
        if (!line.empty() && line.back() == '\r' && nl != std::string::npos) line.pop_back();
        if (!split_whitespace(line, &parts)) {  // `let line = line?;` on a non-UTF-8 line
            delete m;
            set_err(err, errcap, "stream did not contain valid UTF-8");
            return ORA_INVALID_DATA;
        }
        if (line.empty() || line[0] == '#') continue;  // config_loader.rs:22-24
        if (parts.size() != 2) {                       // config_loader.rs:41-43
            delete m;
            set_err(err, errcap, "Invalid merge rule format in line: '" + line +
                                     "'. Expected two numbers separated by space.");
            return ORA_INVALID_DATA;
        }
        uint8_t b1 = 0, b2 = 0;
        std::string e = parse_u8(parts[0], &b1);  // config_loader.rs:27-32
        if (!e.empty()) {
            delete m;
            set_err(err, errcap, "Failed to parse first byte value: " + e + " in line '" + line + "'");
            return ORA_INVALID_DATA;
        }
        e = parse_u8(parts[1], &b2);  // config_loader.rs:33-38
        if (!e.empty()) {
            delete m;
            set_err(err, errcap, "Failed to parse second byte value: " + e + " in line '" + line + "'");
            return ORA_INVALID_DATA;
        }
        // The reference's `vocab_size: u16` overflows after the 65 280th valid line (a debug build
        // panics at `vocab_size += 1`, a release build wraps ids to 0).  That corner is undefined
        // upstream; both this oracle and the CUDA build reject it as InvalidData.
        if (vocab_size > 0xFFFF) {
            delete m;
            set_err(err, errcap, "too many merge rules: token ids exceed u16 (more than 65280 rules)");
            return ORA_INVALID_DATA;
        }
        m->map[pack(b1, b2)] = uint16_t(vocab_size);  // config_loader.rs:39 (later duplicate wins)
        vocab_size += 1;                              // config_loader.rs:40 (every valid line)
    }
    *out = m;
    return ORA_OK;
}

// ---- strategies ---------------------------------------------------------------------------------

namespace {

// tokenizer.rs:56-93.  Returns the BE-serialised token stream as a fresh vector (like the Vec<u8>
// the reference returns).
std::vector<uint8_t> bpe_chunk(const ora_merges *m, const uint8_t *in, size_t n, uint32_t *sweeps) {
    if (sweeps) *sweeps = 0;
    if (n == 0) return {};                                        // tokenizer.rs:57-59
    std::vector<uint16_t> tokens(n);                               // tokenizer.rs:61
    for (size_t i = 0; i < n; ++i) tokens[i] = in[i];
    for (;;) {                                                     // tokenizer.rs:63
        bool merges_found = false;
        std::vector<uint16_t> next;                                // tokenizer.rs:65
        next.reserve(tokens.size());
        size_t i = 0;
        const size_t len = tokens.size();
        while (i < len) {                                          // tokenizer.rs:67
            if (i + 1 < len) {                                     // tokenizer.rs:68
                auto it = m->map.find(pack(tokens[i], tokens[i + 1]));
                if (it != m->map.end()) {                          // tokenizer.rs:69-72
                    next.push_back(it->second);
                    i += 2;
                    merges_found = true;
                    continue;
                }
            }
            next.push_back(tokens[i]);                             // tokenizer.rs:73-80
            i += 1;
        }
        tokens.swap(next);                                         // tokenizer.rs:82
        if (sweeps) *sweeps += 1;
        if (!merges_found) break;                                  // tokenizer.rs:83-85
    }
    std::vector<uint8_t> out;                                      // tokenizer.rs:88-91
    out.reserve(tokens.size() * 2);
    for (uint16_t t : tokens) {
        out.push_back(uint8_t(t >> 8));
        out.push_back(uint8_t(t & 0xff));
    }
    return out;
}

std::vector<uint8_t> basic_chunk(const uint8_t *in, size_t n) {   // tokenizer.rs:108-123
    std::vector<uint8_t> out;
    if (n == 0) return out;
    out.reserve(n * 2);
    for (size_t i = 0; i < n; ++i) {
        out.push_back(0);       // (byte as u16).to_be_bytes() == [0x00, byte]
        out.push_back(in[i]);
    }
    return out;
}

std::vector<uint8_t> passthrough_chunk(const uint8_t *in, size_t n) {  // tokenizer.rs:138-144
    return std::vector<uint8_t>(in, in + n);
}

std::vector<uint8_t> run_strategy(int mode, const ora_merges *m, const uint8_t *in, size_t n) {
    switch (mode) {  // select_strategy, lib.rs:271-282
        case ORA_MODE_PASSTHROUGH: return passthrough_chunk(in, n);
        case ORA_MODE_BPE: return bpe_chunk(m, in, n, nullptr);
        default: return basic_chunk(in, n);
    }
}

int copy_out(const std::vector<uint8_t> &v, uint8_t *out, size_t cap, size_t *out_len) {
    *out_len = v.size();
    if (v.size() > cap) return ORA_CAPACITY;
    if (!v.empty()) std::memcpy(out, v.data(), v.size());
    return ORA_OK;
}

// A sink the ordered writer appends to (memory buffer or buffered file).
struct Sink {
    uint8_t *buf = nullptr;
    size_t cap = 0, len = 0;
    FILE *file = nullptr;
    bool overflow = false, io_error = false;
    void write(const uint8_t *p, size_t n) {
        if (n == 0) return;
        if (file) {
            if (std::fwrite(p, 1, n, file) != n) io_error = true;
            len += n;
            return;
        }
        if (len + n > cap) { overflow = true; len += n; return; }
        std::memcpy(buf + len, p, n);
        len += n;
    }
};

// pipeline.rs:56-131 -- chunk table, <= num_threads chunks in flight, ordered single writer.
void run_pipeline(int mode, const ora_merges *m, const uint8_t *in, size_t n, size_t chunk_size,
                  size_t threads, Sink *sink) {
    if (chunk_size == 0) chunk_size = 1;
    if (threads == 0) threads = 1;
    const size_t n_chunks = (n + chunk_size - 1) / chunk_size;  // slice.chunks(), pipeline.rs:73-81
    if (n_chunks == 0) return;                                  // pipeline.rs:103-105
    threads = std::min(threads, n_chunks);

    std::vector<std::vector<uint8_t>> results(n_chunks);
    std::vector<char> done(n_chunks, 0);
    std::mutex mu;
    std::condition_variable cv_done, cv_slot;
    size_t next_dispatch = 0, next_write = 0, in_flight = 0;
    // In the reference at most `threads` tasks are in flight and the result channel holds
    // 2*threads more (pipeline.rs:68, 86); dispatch stalls beyond that window.
    const size_t window = threads * 3;

    auto worker = [&]() {
        for (;;) {
            size_t id;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_slot.wait(lk, [&] {
                    return next_dispatch >= n_chunks || next_dispatch < next_write + window;
                });
                if (next_dispatch >= n_chunks) return;
                id = next_dispatch++;
                ++in_flight;
            }
            const size_t start = id * chunk_size;
            const size_t len = std::min(chunk_size, n - start);
            std::vector<uint8_t> r = run_strategy(mode, m, in + start, len);  // pipeline.rs:143-144
            {
                std::lock_guard<std::mutex> lk(mu);
                results[id] = std::move(r);
                done[id] = 1;
                --in_flight;
            }
            cv_done.notify_one();
        }
    };
    std::vector<std::thread> pool;
    for (size_t t = 0; t < threads; ++t) pool.emplace_back(worker);
    // Ordered writer: pipeline.rs:153-168.
    while (next_write < n_chunks) {
        std::vector<uint8_t> r;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return done[next_write] != 0; });
            r = std::move(results[next_write]);
        }
        sink->write(r.data(), r.size());
        {
            std::lock_guard<std::mutex> lk(mu);
            ++next_write;
        }
        cv_slot.notify_all();
    }
    for (auto &t : pool) t.join();
}

}  // namespace

extern "C" {

int ora_bpe_process_chunk(const ora_merges *m, const uint8_t *in, size_t n, uint8_t *out,
                          size_t out_cap, size_t *out_len, uint32_t *sweeps) {
    return copy_out(bpe_chunk(m, in, n, sweeps), out, out_cap, out_len);
}
int ora_basic_process_chunk(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                            size_t *out_len) {
    return copy_out(basic_chunk(in, n), out, out_cap, out_len);
}
int ora_passthrough_process_chunk(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                                  size_t *out_len) {
    return copy_out(passthrough_chunk(in, n), out, out_cap, out_len);
}

uint16_t ora_content_type_token(int ct) {  // lib.rs:96-103
    switch (ct) {
        case 0: return 0xFF01;  // Text
        case 1: return 0xFF02;  // Audio
        case 2: return 0xFF03;  // Bin
        default: return 0xFF04; // Video
    }
}

// Inverse of tokenizer.rs:88-91 + lib.rs:284-294 (no reference counterpart; see the header).
int ora_detokenize(const ora_merges *m, const uint8_t *in, size_t n_bytes, int has_content_type,
                   uint8_t *out, size_t out_cap, size_t *out_len) {
    *out_len = 0;
    std::vector<int32_t> inverse(65536, -1);  // id -> left | right << 8
    if (m) {
        for (const auto &kv : m->map) {
            const uint32_t l = kv.first >> 16, r = kv.first & 0xffffu;
            if (l > 255 || r > 255 || kv.second < 256 || inverse[kv.second] >= 0) return ORA_INVALID_INPUT;
            inverse[kv.second] = int32_t(l | (r << 8));
        }
    }
    if (n_bytes & 1) return ORA_INVALID_DATA;
    size_t i = 0;
    if (has_content_type) {
        if (n_bytes < 2) return ORA_INVALID_DATA;
        const uint32_t t = (uint32_t(in[0]) << 8) | in[1];
        if (t < 0xFF01 || t > 0xFF04) return ORA_INVALID_DATA;
        i = 2;
    }
    std::vector<uint8_t> bytes;
    bytes.reserve(n_bytes);
    for (; i < n_bytes; i += 2) {
        const uint32_t t = (uint32_t(in[i]) << 8) | in[i + 1];  // u16::from_be_bytes
        if (t < 256) { bytes.push_back(uint8_t(t)); continue; }
        if (inverse[t] < 0) return ORA_INVALID_DATA;
        bytes.push_back(uint8_t(inverse[t] & 0xff));
        bytes.push_back(uint8_t(inverse[t] >> 8));
    }
    return copy_out(bytes, out, out_cap, out_len);
}

int ora_run_buffer(int mode, const ora_merges *m, const uint8_t *in, size_t n, size_t chunk_size,
                   size_t threads, int content_type_token, uint8_t *out, size_t out_cap,
                   size_t *out_len) {
    Sink sink;
    sink.buf = out;
    sink.cap = out_cap;
    if (content_type_token >= 0) {  // prepend_content_type_token, lib.rs:284-294
        uint8_t be[2] = {uint8_t(content_type_token >> 8), uint8_t(content_type_token & 0xff)};
        sink.write(be, 2);
    }
    run_pipeline(mode, m, in, n, chunk_size, threads, &sink);
    *out_len = sink.len;
    return sink.overflow ? ORA_CAPACITY : ORA_OK;
}

int ora_run_files(int mode, const ora_merges *m, const char *in_path, const char *out_path,
                  size_t chunk_size, size_t threads, int content_type_token, char *err,
                  size_t errcap) {
    int fd = open(in_path, O_RDONLY);  // io_handler.rs:53-54
    if (fd < 0) {
        set_err(err, errcap, std::string(std::strerror(errno)));
        return errno_to_kind(errno);
    }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); set_err(err, errcap, "fstat failed"); return ORA_IO; }
    const size_t n = size_t(st.st_size);
    const uint8_t *in = nullptr;
    void *map = nullptr;
    if (n > 0) {  // memmap2 maps an empty file as an empty slice (io_handler.rs:55)
        map = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (map == MAP_FAILED) { close(fd); set_err(err, errcap, "mmap failed"); return ORA_IO; }
        in = static_cast<const uint8_t *>(map);
    }
    FILE *f = std::fopen(out_path, "wb");  // File::create + BufWriter, io_handler.rs:70-72
    if (!f) {
        int k = errno_to_kind(errno);
        set_err(err, errcap, std::string(std::strerror(errno)));
        if (map) munmap(map, n);
        close(fd);
        return k;
    }
    Sink sink;
    sink.file = f;
    if (content_type_token >= 0) {
        uint8_t be[2] = {uint8_t(content_type_token >> 8), uint8_t(content_type_token & 0xff)};
        sink.write(be, 2);
    }
    run_pipeline(mode, m, in, n, chunk_size, threads, &sink);
    if (std::fclose(f) != 0) sink.io_error = true;  // output_writer.flush(), pipeline.rs:129
    if (map) munmap(map, n);
    close(fd);
    if (sink.io_error) { set_err(err, errcap, "write failed"); return ORA_IO; }
    return ORA_OK;
}

// utils.rs:10-45
int ora_parse_chunk_size(const char *s, size_t *out, char *err, size_t errcap) {
    std::string str(s ? s : "");
    // str::trim(): strip Unicode whitespace at both ends (ASCII subset + NBSP handled bytewise here;
    // size strings are ASCII in every caller).
    auto is_ws = [](unsigned char c) { return c == ' ' || (c >= 0x09 && c <= 0x0D); };
    size_t b = 0, e = str.size();
    while (b < e && is_ws(str[b])) ++b;
    while (e > b && is_ws(str[e - 1])) --e;
    std::string t = str.substr(b, e - b);
    if (t.empty()) { set_err(err, errcap, "Input string is empty"); return ORA_INVALID_INPUT; }  // :12-14
    std::string upper = t;
    for (auto &c : upper) if (c >= 'a' && c <= 'z') c = char(c - 'a' + 'A');  // to_uppercase (ASCII)
    std::string num_part, unit;
    const bool has_unit = upper.size() >= 2 && (upper.compare(upper.size() - 2, 2, "KB") == 0 ||
                                                upper.compare(upper.size() - 2, 2, "MB") == 0);
    if (has_unit) {  // utils.rs:19-20
        num_part = t.substr(0, t.size() - 2);
        unit = upper.substr(upper.size() - 2);
    } else if (std::all_of(upper.begin(), upper.end(), [](char c) { return c >= '0' && c <= '9'; })) {
        num_part = t;  // utils.rs:21-23
    } else {           // utils.rs:24-29
        set_err(err, errcap, "Invalid unit or format: '" + t +
                                 "'. Number must be followed by KB, MB, or be raw bytes.");
        return ORA_INVALID_INPUT;
    }
    if (num_part.empty() && !unit.empty()) {  // utils.rs:31-33
        set_err(err, errcap, "Number part missing for unit '" + t.substr(t.size() - 2) + "'");
        return ORA_INVALID_INPUT;
    }
    // num_part.parse::<usize>(): optional '+', digits only, must fit 64 bits (utils.rs:35-37).
    size_t i = 0;
    if (num_part[0] == '+' && num_part.size() > 1) i = 1;
    unsigned __int128 v = 0;
    bool bad = (i >= num_part.size());
    for (; i < num_part.size() && !bad; ++i) {
        if (num_part[i] < '0' || num_part[i] > '9') { bad = true; break; }
        v = v * 10 + unsigned(num_part[i] - '0');
        if (v > (unsigned __int128)UINT64_MAX) bad = true;
    }
    if (bad) { set_err(err, errcap, "Invalid number: '" + num_part + "'"); return ORA_INVALID_INPUT; }
    unsigned __int128 r = v;
    if (unit == "KB") r = v * 1024;                 // utils.rs:40
    else if (unit == "MB") r = v * 1024 * 1024;     // utils.rs:41
    if (r > (unsigned __int128)UINT64_MAX) {        // overflow panics/wraps upstream; rejected here
        set_err(err, errcap, "Invalid number: '" + num_part + "'");
        return ORA_INVALID_INPUT;
    }
    *out = size_t(r);
    return ORA_OK;
}

// chunking.rs:18-62
size_t ora_effective_chunk_size(int has_cli, size_t cli_size, size_t threads, unsigned memcap,
                                uint64_t total_ram_bytes) {
    const size_t DEFAULT_MIN = 1024 * 1024, DEFAULT_MAX = 16 * 1024 * 1024;    // chunking.rs:18-19
    const size_t ABS_MIN = 256 * 1024, ABS_MAX = 128 * 1024 * 1024;            // chunking.rs:20-21
    auto clamp = [](size_t v, size_t lo, size_t hi) { return v < lo ? lo : (v > hi ? hi : v); };
    if (has_cli) return clamp(cli_size, ABS_MIN, ABS_MAX);                     // chunking.rs:27-30
    const double usable_f = double(total_ram_bytes) * (double(memcap) / 100.0);  // chunking.rs:41-42
    // Rust `as u64` saturates; the product is far below 2^64 for any real RAM size.
    const uint64_t usable = usable_f >= 18446744073709551615.0 ? UINT64_MAX : uint64_t(usable_f);
    if (threads == 0) threads = 1;  // num_threads is never 0 upstream (utils.rs:81-86)
    const uint64_t per_thread = usable / uint64_t(threads);                    // chunking.rs:50
    const size_t calculated = size_t(per_thread / 4);                          // chunking.rs:56
    return clamp(clamp(calculated, DEFAULT_MIN, DEFAULT_MAX), ABS_MIN, ABS_MAX);  // chunking.rs:59-61
}

// utils.rs:79-97
size_t ora_determine_thread_count(int has_override, size_t override_val, size_t logical_cpus) {
    if (has_override) return override_val == 0 ? 1 : override_val;
    return logical_cpus > 0 ? logical_cpus : 1;
}

}  // extern "C"
