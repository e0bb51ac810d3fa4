// host_config.cpp -- see host_config.h.  Semantics follow the cited reference lines; the code is
// this build's own (streaming line reader, table-driven integer parser).
#include "host_config.h"

#include "../../include/blt_cuda.h"

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sched.h>
#include <thread>

namespace blth {
namespace {

Error make_error(int code, std::string msg) {
    Error e;
    e.code = code;
    e.msg = std::move(msg);
    return e;
}

// ---- UTF-8 scanning (BufRead::lines yields String, so a non-UTF-8 line is an InvalidData error) ----
// Returns the scalar starting at s[i] and advances i, or -1 for an ill-formed sequence.
long next_scalar(const std::string &s, size_t &i) {
    auto cont = [&](size_t k) { return k < s.size() && (static_cast<unsigned char>(s[k]) & 0xC0) == 0x80; };
    const unsigned char b0 = static_cast<unsigned char>(s[i]);
    if (b0 < 0x80) { ++i; return b0; }
    int need;
    long cp, min_cp;
    if (b0 >= 0xC2 && b0 <= 0xDF) { need = 1; cp = b0 & 0x1F; min_cp = 0x80; }
    else if ((b0 & 0xF0) == 0xE0) { need = 2; cp = b0 & 0x0F; min_cp = 0x800; }
    else if (b0 >= 0xF0 && b0 <= 0xF4) { need = 3; cp = b0 & 0x07; min_cp = 0x10000; }
    else return -1;
    for (int k = 1; k <= need; ++k) {
        if (!cont(i + k)) return -1;
        cp = (cp << 6) | (static_cast<unsigned char>(s[i + k]) & 0x3F);
    }
    if (cp < min_cp || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return -1;
    i += size_t(need) + 1;
    return cp;
}

// char::is_whitespace (Unicode White_Space), the separator set of str::split_whitespace.
bool is_white_space(long c) {
    switch (c) {
        case 0x20: case 0x85: case 0xA0: case 0x1680: case 0x2028: case 0x2029: case 0x202F: case 0x205F:
        case 0x3000:
            return true;
        default:
            return (c >= 0x09 && c <= 0x0D) || (c >= 0x2000 && c <= 0x200A);
    }
}

// Splits like str::split_whitespace; false if the line is not UTF-8.
bool fields_of(const std::string &line, std::vector<std::string> *fields) {
    fields->clear();
    std::string cur;
    size_t i = 0;
    while (i < line.size()) {
        const size_t at = i;
        const long cp = next_scalar(line, i);
        if (cp < 0) return false;
        if (is_white_space(cp)) {
            if (!cur.empty()) { fields->push_back(cur); cur.clear(); }
        } else {
            cur.append(line, at, i - at);
        }
    }
    if (!cur.empty()) fields->push_back(cur);
    return true;
}

// <u8 as FromStr>::from_str: [+]digits, 0..=255.  Returns nullptr on success, else the
// ParseIntError Display string.
const char *parse_byte(const std::string &f, uint8_t *out) {
    static const char *kEmpty = "cannot parse integer from empty string";
    static const char *kDigit = "invalid digit found in string";
    static const char *kBig = "number too large to fit in target type";
    if (f.empty()) return kEmpty;
    size_t k = 0;
    if (f[0] == '+') {
        if (f.size() == 1) return kDigit;
        k = 1;
    }
    unsigned acc = 0;
    for (; k < f.size(); ++k) {
        const char c = f[k];
        if (c < '0' || c > '9') return kDigit;  // '-' lands here: unsigned types take no minus sign
        acc = acc * 10u + unsigned(c - '0');
        if (acc > 0xFFu) return kBig;            // reported where the overflow happens
    }
    *out = uint8_t(acc);
    return nullptr;
}

int kind_of_errno(int e) { return e == ENOENT ? BLT_ERR_NOT_FOUND : BLT_ERR_IO; }

}  // namespace

MergeList dedup_rules(const std::vector<MergeRule> &rules) {
    std::map<uint32_t, uint16_t> last;  // ordered => output sorted by (left,right)
    for (const MergeRule &r : rules) last[(uint32_t(r.left) << 16) | r.right] = r.value;
    MergeList out;
    out.reserve(last.size());
    for (const auto &kv : last) out.push_back(MergeRule{uint16_t(kv.first >> 16), uint16_t(kv.first & 0xFFFF), kv.second});
    return out;
}

Error load_merges_file(const std::string &path, MergeList *out) {
    out->clear();
    errno = 0;
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) {  // File::open(path)?  (config_loader.rs:15)
        const int e = errno ? errno : ENOENT;
        return make_error(kind_of_errno(e), std::string(std::strerror(e)) + " (os error " + std::to_string(e) + ")");
    }
    std::vector<MergeRule> rules;
    std::vector<std::string> fields;
    std::string line;
    uint32_t next_id = 256;  // config_loader.rs:18
    while (std::getline(f, line)) {
        // BufRead::lines drops "\n" and, when present before it, "\r".  getline already consumed the
        // "\n"; eof() tells a final unterminated line apart, whose "\r" (if any) is data.
        if (!f.eof() && !line.empty() && line.back() == '\r') line.pop_back();
        if (!fields_of(line, &fields))
            return make_error(BLT_ERR_INVALID_DATA, "stream did not contain valid UTF-8");
        if (line.empty() || line.front() == '#') continue;  // config_loader.rs:22-24
        if (fields.size() != 2)                             // config_loader.rs:41-43
            return make_error(BLT_ERR_INVALID_DATA, "Invalid merge rule format in line: '" + line +
                                                        "'. Expected two numbers separated by space.");
        uint8_t a = 0, b = 0;
        if (const char *why = parse_byte(fields[0], &a))    // config_loader.rs:27-32
            return make_error(BLT_ERR_INVALID_DATA,
                              std::string("Failed to parse first byte value: ") + why + " in line '" + line + "'");
        if (const char *why = parse_byte(fields[1], &b))    // config_loader.rs:33-38
            return make_error(BLT_ERR_INVALID_DATA,
                              std::string("Failed to parse second byte value: ") + why + " in line '" + line + "'");
        // The reference counts ids in a u16 (config_loader.rs:18,40): the 65 281st valid line would
        // overflow it (debug: panic; release: wraps to 0).  Undefined upstream; rejected here.
        if (next_id > 0xFFFFu)
            return make_error(BLT_ERR_INVALID_DATA, "too many merge rules: token ids exceed u16 (more than 65280 rules)");
        rules.push_back(MergeRule{a, b, uint16_t(next_id)});
        ++next_id;  // every valid line consumes an id, duplicates included (config_loader.rs:40)
    }
    if (f.bad()) return make_error(BLT_ERR_IO, "read error on merges file");
    *out = dedup_rules(rules);
    return Error{};
}

Error parse_chunk_size(const std::string &raw, size_t *out) {
    // utils.rs:11 trim()
    size_t lo = 0, hi = raw.size();
    while (lo < hi && std::isspace(static_cast<unsigned char>(raw[lo]))) ++lo;
    while (hi > lo && std::isspace(static_cast<unsigned char>(raw[hi - 1]))) --hi;
    const std::string s = raw.substr(lo, hi - lo);
    if (s.empty()) return make_error(BLT_ERR_INVALID_INPUT, "Input string is empty");  // utils.rs:12-14
    unsigned shift = 0;
    std::string digits = s, unit;
    if (s.size() >= 2) {
        const char u0 = char(std::toupper(static_cast<unsigned char>(s[s.size() - 2])));
        const char u1 = char(std::toupper(static_cast<unsigned char>(s[s.size() - 1])));
        if (u1 == 'B' && (u0 == 'K' || u0 == 'M')) {  // utils.rs:19-20
            shift = (u0 == 'K') ? 10 : 20;
            unit = s.substr(s.size() - 2);
            digits = s.substr(0, s.size() - 2);
        }
    }
    if (unit.empty() && !std::all_of(s.begin(), s.end(), [](char c) { return c >= '0' && c <= '9'; }))
        return make_error(BLT_ERR_INVALID_INPUT,  // utils.rs:24-29
                          "Invalid unit or format: '" + s + "'. Number must be followed by KB, MB, or be raw bytes.");
    if (digits.empty())  // utils.rs:31-33
        return make_error(BLT_ERR_INVALID_INPUT, "Number part missing for unit '" + unit + "'");
    // utils.rs:35-37  usize::from_str: [+]digits, must fit 64 bits
    size_t k = (digits[0] == '+' && digits.size() > 1) ? 1 : 0;
    uint64_t acc = 0;
    bool ok = true;
    for (; k < digits.size() && ok; ++k) {
        const char c = digits[k];
        if (c < '0' || c > '9') { ok = false; break; }
        if (acc > (UINT64_MAX - uint64_t(c - '0')) / 10) { ok = false; break; }
        acc = acc * 10 + uint64_t(c - '0');
    }
    if (!ok) return make_error(BLT_ERR_INVALID_INPUT, "Invalid number: '" + digits + "'");
    if (shift && (acc >> (64 - shift)) != 0)  // num * 1024 overflow: panic/wrap upstream, rejected here
        return make_error(BLT_ERR_INVALID_INPUT, "Invalid number: '" + digits + "'");
    *out = size_t(acc << shift);  // utils.rs:39-43
    return Error{};
}

uint64_t host_total_ram() {
    std::ifstream f("/proc/meminfo");
    std::string key;
    uint64_t kb = 0;
    while (f >> key) {
        if (key == "MemTotal:") { f >> kb; break; }
        std::getline(f, key);
    }
    return kb * 1024ull;
}

size_t host_logical_cpus() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof set, &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return size_t(n);
    }
    const unsigned h = std::thread::hardware_concurrency();
    return h ? h : 1;
}

size_t effective_chunk_size(bool has_cli, size_t cli, size_t threads, unsigned memcap, uint64_t total_ram) {
    constexpr size_t kMiB = 1024 * 1024;
    const size_t abs_lo = 256 * 1024, abs_hi = 128 * kMiB;  // chunking.rs:20-21
    if (has_cli) return std::min(std::max(cli, abs_lo), abs_hi);  // chunking.rs:27-30
    if (total_ram == 0) total_ram = host_total_ram();             // chunking.rs:33-37
    const double usable_f = double(total_ram) * (double(memcap) / 100.0);  // chunking.rs:41-42
    const uint64_t usable = usable_f >= 1.8446744073709552e19 ? UINT64_MAX : uint64_t(usable_f);
    const uint64_t per_thread = usable / uint64_t(threads ? threads : 1);  // chunking.rs:50
    size_t c = size_t(per_thread / 4);                                     // chunking.rs:56
    c = std::min(std::max(c, 1 * kMiB), 16 * kMiB);                        // chunking.rs:59-60
    return std::min(std::max(c, abs_lo), abs_hi);                          // chunking.rs:61
}

size_t determine_thread_count(bool has_override, size_t value) {
    if (has_override) return value ? value : 1;  // utils.rs:80-87
    return host_logical_cpus();                  // utils.rs:88-95
}

}  // namespace blth
