// synth.cpp -- deterministic synthetic workloads for the five BASELINE.json configs.
// Host-only helper library (libblt_synth.so) shared by tests and bench.py so that the oracle, the
// CPU baseline and the GPU path all see bit-identical inputs.  Not part of the tokenizer itself.
//
// PRNG: splitmix64.  Because its state advances by a constant, output k of a stream is a pure
// function of (seed, k), which makes every generator below block-parallel and reproducible.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct SplitMix {
    uint64_t s;
    explicit SplitMix(uint64_t seed) : s(seed) {}
    uint64_t next() {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};

uint64_t mix(uint64_t a, uint64_t b) {
    SplitMix m(a ^ (b * 0xD1342543DE82EF95ull));
    return m.next();
}

// ---- English-like text: Zipf(s=1) over a fixed 4 096-word lexicon ---------------------------------
constexpr int kLexicon = 4096;
constexpr size_t kTextBlock = 1u << 20;  // generation unit: blocks are independent

struct Lexicon {
    char words[kLexicon][12];
    uint8_t len[kLexicon];
    uint64_t cum[kLexicon];  // cumulative integer Zipf weights floor(2^32 / rank)
    uint64_t total;
    Lexicon() {
        // letters by English unigram frequency (per mille, sums to 1000)
        static const struct { char c; int w; } freq[26] = {
            {'e', 127}, {'t', 91}, {'a', 82}, {'o', 75}, {'i', 70}, {'n', 67}, {'s', 63}, {'h', 61}, {'r', 60},
            {'d', 43},  {'l', 40}, {'c', 28}, {'u', 28}, {'m', 24}, {'w', 24}, {'f', 22}, {'g', 20}, {'y', 20},
            {'p', 19},  {'b', 15}, {'v', 10}, {'k', 8},  {'j', 1},  {'x', 1},  {'q', 1},  {'z', 0}};
        // word length 1..12, weights in percent
        static const int len_w[12] = {3, 12, 20, 18, 13, 10, 8, 6, 4, 3, 2, 1};
        SplitMix rng(0xB171E81ull);
        for (int w = 0; w < kLexicon; ++w) {
            int r = int(rng.next() % 100), L = 0;
            while (r >= len_w[L]) { r -= len_w[L]; ++L; }
            len[w] = uint8_t(L + 1);
            for (int k = 0; k <= L; ++k) {
                int f = int(rng.next() % 1000), c = 0;
                while (f >= freq[c].w) { f -= freq[c].w; ++c; }
                words[w][k] = freq[c].c;
            }
        }
        uint64_t acc = 0;
        for (int w = 0; w < kLexicon; ++w) {
            acc += (1ull << 32) / uint64_t(w + 1);
            cum[w] = acc;
        }
        total = acc;
    }
    int draw(SplitMix &rng) const {
        const uint64_t u = rng.next() % total;
        return int(std::upper_bound(cum, cum + kLexicon, u) - cum);
    }
};

const Lexicon &lexicon() {
    static const Lexicon lx;
    return lx;
}

void text_block(uint8_t *out, size_t len, uint64_t seed, uint64_t block) {
    const Lexicon &lx = lexicon();
    SplitMix rng(mix(seed, block));
    size_t pos = 0;
    while (pos < len) {
        const int w = lx.draw(rng);
        for (int k = 0; k < lx.len[w] && pos < len; ++k) out[pos++] = uint8_t(lx.words[w][k]);
        const int r = int(rng.next() % 1000);
        const char *sep = r < 850 ? " " : r < 900 ? ", " : r < 960 ? ". " : "\n";
        for (const char *p = sep; *p && pos < len; ++p) out[pos++] = uint8_t(*p);
    }
}

void fill_period(uint8_t *out, size_t len, const char *pat) {
    const size_t p = std::strlen(pat);
    for (size_t i = 0; i < len; ++i) out[i] = uint8_t(pat[i % p]);
}

// ---- mixed corpus (round 2): what a crawl or an archive looks like to a byte-level tokenizer ---------------
// Segments of five kinds follow each other inside independent 1 MiB blocks (30 % English-like text from the lexicon
// above, 12 % code/markup, 8 % UTF-8 script, 10 % base64, 40 % binary):
// code/markup is ~90 printable ASCII symbols (Zipf 0.7), the script text is lead D0/D1 + continuation bytes,
// base64 blobs are 64 uniform symbols and binary data uniform bytes.  All 256 byte values and all
// 65 536 pairs occur, so a 32 768-rule table trained on it is a STRICT subset of the observed pairs: the sweep
// meets non-rule pairs all the time (T_out / N_in ~ 0.57) and the carry / compaction machinery really runs.
void mixed_block(uint8_t *out, size_t len, uint64_t seed, uint64_t block) {
    const Lexicon &lx = lexicon();
    static const char code_syms[] = " etaoinsrhldcumfpgwybvkxjqz(){};=.,_\"'<>/-+*&|![]:#0123456789ETAOINSRHLDCUMFPGWYBVKXJQZ\t\n@$%^~?`\\";
    constexpr int n_code = int(sizeof code_syms) - 1;
    static uint64_t code_cum[n_code];
    static uint64_t code_total = [] {
        uint64_t acc = 0;
        for (int i = 0; i < n_code; ++i) {
            // Zipf with exponent 0.7: weight ~ 1 / (i+1)^0.7, in integers
            double w = 1.0;
            for (int k = 0; k < 7; ++k) w *= double(i + 1);
            double root = 1.0;  // (i+1)^0.7 = exp(0.7 ln(i+1)), by 40 Newton steps on x^10 = (i+1)^7
            for (int it = 0; it < 60; ++it) {
                double x9 = 1.0;
                for (int k = 0; k < 9; ++k) x9 *= root;
                root = root - (x9 * root - w) / (10.0 * x9);
            }
            acc += uint64_t(4294967296.0 / root);
            code_cum[i] = acc;
        }
        return acc;
    }();
    static const char b64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    SplitMix rng(mix(seed ^ 0x6D69786564ull, block));
    size_t pos = 0;
    while (pos < len) {
        const uint64_t r = rng.next();
        const int kind_r = int(r % 100);
        size_t seg = size_t(256 + ((r >> 8) % 32768));
        seg = std::min(seg, len - pos);
        const size_t end = pos + seg;
        if (kind_r < 30) {  // English-like text
            while (pos < end) {
                const int w = lx.draw(rng);
                for (int k = 0; k < lx.len[w] && pos < end; ++k) out[pos++] = uint8_t(lx.words[w][k]);
                const int q = int(rng.next() % 1000);
                const char *sep = q < 850 ? " " : q < 900 ? ", " : q < 960 ? ". " : "\n";
                for (const char *p = sep; *p && pos < end; ++p) out[pos++] = uint8_t(*p);
            }
        } else if (kind_r < 42) {  // code / markup
            while (pos < end) {
                const uint64_t u = rng.next() % code_total;
                out[pos++] = uint8_t(code_syms[std::upper_bound(code_cum, code_cum + n_code, u) - code_cum]);
            }
        } else if (kind_r < 50) {  // UTF-8 two-byte script text with spaces
            while (pos < end) {
                uint64_t v = rng.next();
                for (int k = 0; k < 6 && pos + 1 < end; ++k, v >>= 10) {
                    if ((v & 7) == 0) { out[pos++] = ' '; continue; }
                    out[pos++] = uint8_t(0xD0 + ((v >> 3) & 1));
                    out[pos++] = uint8_t(0x80 + ((v >> 4) & 63));
                }
                if (pos + 1 == end) out[pos++] = ' ';
            }
        } else if (kind_r < 60) {  // base64
            while (pos < end) {
                uint64_t v = rng.next();
                for (int k = 0; k < 10 && pos < end; ++k, v >>= 6) out[pos++] = uint8_t(b64[v & 63]);
            }
        } else {  // binary
            while (pos < end) {
                uint64_t v = rng.next();
                for (int k = 0; k < 8 && pos < end; ++k, v >>= 8) out[pos++] = uint8_t(v);
            }
        }
    }
}

}  // namespace

extern "C" {

// Config 1: i.i.d. uniform bytes: little-endian words of the splitmix64 stream.
void blt_synth_random(uint8_t *out, size_t n, uint64_t seed) {
    SplitMix rng(seed);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const uint64_t v = rng.next();
        std::memcpy(out + i, &v, 8);
    }
    if (i < n) {
        const uint64_t v = rng.next();
        std::memcpy(out + i, &v, n - i);
    }
}

// Configs 2, 3, 5: English-like text, independent 1 MiB blocks.
void blt_synth_text(uint8_t *out, size_t n, uint64_t seed, int threads) {
    const size_t blocks = (n + kTextBlock - 1) / kTextBlock;
    if (threads < 1) threads = 1;
    lexicon();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([=]() {
            for (size_t b = size_t(t); b < blocks; b += size_t(threads)) {
                const size_t off = b * kTextBlock;
                text_block(out + off, std::min(kTextBlock, n - off), seed, b);
            }
        });
    }
    for (auto &th : pool) th.join();
}

// Mixed corpus (see mixed_block): independent 1 MiB blocks.
void blt_synth_mixed(uint8_t *out, size_t n, uint64_t seed, int threads) {
    const size_t blocks = (n + kTextBlock - 1) / kTextBlock;
    if (threads < 1) threads = 1;
    lexicon();
    mixed_block(nullptr, 0, seed, 0);  // builds the static tables before the threads start
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([=]() {
            for (size_t b = size_t(t); b < blocks; b += size_t(threads)) {
                const size_t off = b * kTextBlock;
                mixed_block(out + off, std::min(kTextBlock, n - off), seed, b);
            }
        });
    }
    for (auto &th : pool) th.join();
}

// Config 4: adversarial input over the alphabet {a,b,c,d}: single-byte runs of awkward lengths
// (including 2^20 +- 1 and 16 MiB + 1, which straddles a chunk wall), periodic blocks whose candidate
// merges overlap (abab.., aab.., abba..), and random 4-symbol text.
void blt_synth_adversarial(uint8_t *out, size_t n, uint64_t seed) {
    static const size_t run_len[] = {1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65, 1023, 1024, 1025,
                                     (1u << 20) - 1, (1u << 20), (1u << 20) + 1, (16u << 20) + 1};
    static const char *periodic[] = {"ab", "aab", "abba", "ba", "abc", "aabb", "abab" "b", "dcba", "aaab"};
    SplitMix rng(seed);
    size_t pos = 0;
    while (pos < n) {
        const uint64_t r = rng.next();
        const int kind = int(r % 10);
        size_t len;
        if (kind < 4) {  // run of one byte
            len = run_len[(r >> 8) % (sizeof run_len / sizeof run_len[0])];
            len = std::min(len, n - pos);
            std::memset(out + pos, 'a' + int((r >> 20) & 3), len);
        } else if (kind < 7) {  // periodic block
            len = std::min(size_t(1 + ((r >> 8) % 100000)), n - pos);
            fill_period(out + pos, len, periodic[(r >> 32) % (sizeof periodic / sizeof periodic[0])]);
        } else {  // random text over 4 symbols
            len = std::min(size_t(1 + ((r >> 8) % 65536)), n - pos);
            for (size_t i = 0; i < len; i += 32) {
                uint64_t v = rng.next();
                for (size_t k = 0; k < 32 && i + k < len; ++k, v >>= 2) out[pos + i + k] = uint8_t('a' + (v & 3));
            }
        }
        pos += len;
    }
}

// merges table for configs 2, 3, 5: adjacent byte pairs of data[0..n_sample) by descending count,
// ties by b0*256+b1 ascending; if fewer than n_rules pairs occur, never-observed pairs follow in
// ascending b0*256+b1.  Writes n_rules (<= 65 280) pairs; returns the number of OBSERVED pairs used.
size_t blt_synth_merges(const uint8_t *data, size_t n_sample, size_t n_rules, uint8_t *left, uint8_t *right) {
    std::vector<uint64_t> hist(65536, 0);
    for (size_t i = 0; i + 1 < n_sample; ++i) hist[(size_t(data[i]) << 8) | data[i + 1]]++;
    std::vector<uint32_t> keys;
    for (uint32_t k = 0; k < 65536; ++k)
        if (hist[k]) keys.push_back(k);
    std::sort(keys.begin(), keys.end(), [&](uint32_t a, uint32_t b) {
        return hist[a] != hist[b] ? hist[a] > hist[b] : a < b;
    });
    size_t w = 0, observed = 0;
    for (size_t i = 0; i < keys.size() && w < n_rules; ++i, ++w, ++observed) {
        left[w] = uint8_t(keys[i] >> 8);
        right[w] = uint8_t(keys[i] & 0xff);
    }
    for (uint32_t k = 0; k < 65536 && w < n_rules; ++k) {
        if (hist[k]) continue;
        left[w] = uint8_t(k >> 8);
        right[w] = uint8_t(k & 0xff);
        ++w;
    }
    return observed;
}

}  // extern "C"
