// UNTESTED: there is no Rust toolchain in the image this repository is built and tested in (no rustc, cargo or
// maturin).  Everything this file calls is exercised through the same C ABI (include/blt_cuda.h) by the C++ CLI,
// the ctypes binding and tests/.  See INTEGRATION.md.
// Replacement for select_strategy (blt_core/src/lib.rs:271-282): the same precedence (passthrough > bpe > basic),
// CUDA strategies behind the same trait object.
fn select_strategy(config: &CoreConfig) -> Arc<dyn TokenizationStrategy> {
    let dev = 0;
    if config.passthrough_mode { Arc::new(cuda_strategy::CudaStrategy::passthrough(dev).expect("cuda")) }
    else if let Some(ref m) = config.bpe_data { Arc::new(cuda_strategy::CudaStrategy::bpe(dev, m).expect("cuda")) }
    else { Arc::new(cuda_strategy::CudaStrategy::basic(dev).expect("cuda")) }
}

// Optional: the whole pipeline in one native call instead of pipeline::run (blt_core/src/lib.rs:245-267).
#[repr(C)]
pub struct blt_core_config {
    input: *const c_char, output: *const c_char, merges_file: *const c_char, content_type: c_int,
    has_threads: c_int, threads: usize, chunk_size: *const c_char, has_memcap: c_int, memcap: u32,
    passthrough: c_int, num_gpus: c_int,
}
extern "C" { fn blt_run_tokenizer(cfg: *const blt_core_config) -> c_int; }
