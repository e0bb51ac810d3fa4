// cli_main.cpp -- the `blt` command line of the reference (src/main.rs:8-106) over libblt_cuda.so.
// Same flags, same defaults, same exit behaviour; one addition: --gpus N (default: 1).
#include "../../include/blt_cuda.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

namespace {

const char *kHelp =
    "Usage: blt [OPTIONS]\n"
    "\n"
    "Options:\n"
    "  -i, --input <FILE>        Input file path (or - for stdin)\n"
    "  -o, --output <FILE>       Output file path (or - for stdout)\n"
    "      --merges <FILE>       BPE merges file for advanced tokenization\n"
    "      --passthrough         Use passthrough mode (copy file without tokenization)\n"
    "      --type <TYPE>         Prepend content-type token [possible values: text, audio, bin, video]\n"
    "      --threads <NUM>       Override worker count (default: auto based on cores)\n"
    "      --memcap <PERCENT>    Max RAM usage fraction (e.g., 70 for 70%)\n"
    "      --chunksize <SIZE>    Min/Max chunk size (e.g. 4MB, 256KB).\n"
    "      --detokenize          Inverse: INPUT holds tokens, OUTPUT receives the bytes (--type: a content-type token is present)\n"
    "      --gpus <NUM>          GPUs to shard chunks over (default: 1; every extra GPU costs ~1 s of CUDA start-up)\n"
    "  -h, --help                Print help\n"
    "  -V, --version             Print version\n";

[[noreturn]] void usage_error(const std::string &msg) {  // clap: message on stderr, exit code 2
    std::fprintf(stderr, "error: %s\n\nUsage: blt [OPTIONS]\n\nFor more information, try '--help'.\n", msg.c_str());
    std::exit(2);
}

bool parse_unsigned(const char *s, unsigned long long max, unsigned long long *out) {
    if (!*s) return false;
    const char *p = s;
    if (*p == '+' && p[1]) ++p;
    unsigned long long v = 0;
    for (; *p; ++p) {
        if (*p < '0' || *p > '9') return false;
        if (v > (max - unsigned(*p - '0')) / 10) return false;
        v = v * 10 + unsigned(*p - '0');
    }
    *out = v;
    return true;
}

const char *kind_name(int code) {
    switch (code) {
        case BLT_ERR_NOT_FOUND: return "NotFound";
        case BLT_ERR_INVALID_INPUT: return "InvalidInput";
        case BLT_ERR_INVALID_DATA: return "InvalidData";
        default: return "Other";
    }
}

}  // namespace

int main(int argc, char **argv) {
    std::string input, output, merges, type, chunksize;
    bool has_input = false, has_output = false, has_merges = false, has_chunk = false, passthrough = false;
    bool has_threads = false, has_memcap = false, detokenize = false;
    unsigned long long threads = 0, memcap = 0, gpus = 0;
    int content_type = BLT_CONTENT_NONE;

    std::vector<std::string> args(argv + 1, argv + argc);
    for (size_t i = 0; i < args.size(); ++i) {
        std::string a = args[i], val;
        bool has_val = false;
        if (a.rfind("--", 0) == 0) {
            const size_t eq = a.find('=');
            if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; }
        } else if (a.size() > 2 && a[0] == '-' && (a[1] == 'i' || a[1] == 'o')) {  // -iFILE / -i=FILE
            val = a.substr(a[2] == '=' ? 3 : 2);
            a = a.substr(0, 2);
            has_val = true;
        }
        auto need = [&](const char *what) -> std::string {
            if (has_val) return val;
            if (i + 1 >= args.size()) usage_error(std::string("a value is required for '") + what + "' but none was supplied");
            return args[++i];
        };
        if (a == "-h" || a == "--help") { std::fputs(kHelp, stdout); return 0; }
        if (a == "-V" || a == "--version") { std::printf("blt %s\n", blt_version()); return 0; }
        if (a == "-i" || a == "--input") { input = need("--input <FILE>"); has_input = true; }
        else if (a == "-o" || a == "--output") { output = need("--output <FILE>"); has_output = true; }
        else if (a == "--merges") { merges = need("--merges <FILE>"); has_merges = true; }
        else if (a == "--passthrough") { if (has_val) usage_error("unexpected value for '--passthrough'"); passthrough = true; }
        else if (a == "--type") {
            type = need("--type <TYPE>");
            if (type == "text") content_type = BLT_CONTENT_TEXT;
            else if (type == "audio") content_type = BLT_CONTENT_AUDIO;
            else if (type == "bin") content_type = BLT_CONTENT_BIN;
            else if (type == "video") content_type = BLT_CONTENT_VIDEO;
            else usage_error("invalid value '" + type + "' for '--type <TYPE>'\n  [possible values: text, audio, bin, video]");
        } else if (a == "--threads") {
            const std::string v = need("--threads <NUM>");
            if (!parse_unsigned(v.c_str(), ~0ull, &threads)) usage_error("invalid value '" + v + "' for '--threads <NUM>'");
            has_threads = true;
        } else if (a == "--memcap") {
            const std::string v = need("--memcap <PERCENT>");
            if (!parse_unsigned(v.c_str(), 255, &memcap)) usage_error("invalid value '" + v + "' for '--memcap <PERCENT>'");
            has_memcap = true;
        } else if (a == "--chunksize") { chunksize = need("--chunksize <SIZE>"); has_chunk = true; }
        else if (a == "--detokenize") { if (has_val) usage_error("unexpected value for '--detokenize'"); detokenize = true; }
        else if (a == "--gpus") {
            const std::string v = need("--gpus <NUM>");
            if (!parse_unsigned(v.c_str(), 1024, &gpus)) usage_error("invalid value '" + v + "' for '--gpus <NUM>'");
        } else usage_error("unexpected argument '" + args[i] + "' found");
    }

    // CoreConfig::new_from_cli failures make `main` return Err: Rust prints `Error: {e:?}`, exit 1.
    if (has_chunk) {
        size_t tmp = 0;
        if (blt_parse_chunk_size(chunksize.c_str(), &tmp) != BLT_OK) {
            std::fprintf(stderr, "Error: Custom { kind: InvalidInput, error: \"%s\" }\n", blt_last_error());
            return 1;
        }
    }
    if (has_merges) {
        size_t n = 0;
        if (blt_load_bpe_merges(merges.c_str(), nullptr, nullptr, nullptr, 0, &n) != BLT_OK) {
            std::fprintf(stderr, "Error: Custom { kind: InvalidInput, error: \"Failed to load BPE merges: %s\" }\n",
                         blt_last_error());
            return 1;
        }
    }

    blt_core_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.input = has_input ? input.c_str() : nullptr;
    cfg.output = has_output ? output.c_str() : nullptr;
    cfg.merges_file = has_merges ? merges.c_str() : nullptr;
    cfg.content_type = content_type;
    cfg.has_threads = has_threads;
    cfg.threads = size_t(threads);
    cfg.chunk_size = has_chunk ? chunksize.c_str() : nullptr;
    cfg.has_memcap = has_memcap;
    cfg.memcap = unsigned(memcap);
    cfg.passthrough = passthrough;
    cfg.num_gpus = int(gpus);
    // CUDA start-up time grows with the number of devices the driver has to open (about a second each on an
    // 8-GPU box): unless the user chose the devices, expose only the ones this run will use.
    if (getenv("CUDA_VISIBLE_DEVICES") == nullptr) {
        std::string vis;
        for (unsigned long long g = 0; g < (gpus ? gpus : 1); ++g) vis += (g ? "," : "") + std::to_string(g);
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 0);
    }
    const int rc = detokenize ? blt_run_detokenizer(&cfg) : blt_run_tokenizer(&cfg);
    if (rc != BLT_OK) {  // main.rs:100-103
        std::fprintf(stderr, "Error running tokenizer: %s\n", blt_last_error());
        (void)kind_name;
        std::fflush(nullptr);
        _exit(1);
    }
    // Everything is written and closed; leave without the CUDA runtime's exit-time teardown of the primary
    // context (0.2-0.3 s that a one-shot command line tool has no use for).
    std::fflush(nullptr);
    _exit(0);
}
