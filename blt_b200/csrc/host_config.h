// host_config.h -- host-side configuration semantics of blt_core, re-implemented for the CUDA build:
// merges-file loader, size-string grammar, chunk sizing, thread count.  No device code.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace blth {

// Error carrier: code is a blt_status value (include/blt_cuda.h), msg the io::Error text.
struct Error {
    int code = 0;
    std::string msg;
    explicit operator bool() const { return code != 0; }
};

struct MergeRule {
    uint16_t left, right, value;
};

// The final content of a HashMap<(u16,u16),u16>, sorted by (left,right), keys unique.
using MergeList = std::vector<MergeRule>;

// load_bpe_merges_from_path (blt_core/src/config_loader.rs:14-46).
Error load_merges_file(const std::string &path, MergeList *out);
// HashMap::insert over the given rules in order: a later duplicate key wins (config_loader.rs:39).
MergeList dedup_rules(const std::vector<MergeRule> &rules_in_insert_order);

// parse_chunk_size_str (blt_core/src/utils.rs:10-45).
Error parse_chunk_size(const std::string &s, size_t *out);
// get_effective_chunk_size (blt_core/src/chunking.rs:26-62); total_ram == 0 probes the host.
size_t effective_chunk_size(bool has_cli, size_t cli, size_t threads, unsigned memcap, uint64_t total_ram);
// determine_thread_count (blt_core/src/utils.rs:79-97).
size_t determine_thread_count(bool has_override, size_t value);
// sysinfo total_memory(): MemTotal of /proc/meminfo, in bytes.
uint64_t host_total_ram();
// num_cpus::get(): CPUs this process may run on.
size_t host_logical_cpus();

}  // namespace blth
