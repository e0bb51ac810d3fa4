// UNTESTED: there is no Rust toolchain in the image this repository is built and tested in (no rustc, cargo or
// maturin).  Everything this file calls is exercised through the same C ABI (include/blt_cuda.h) by the C++ CLI,
// the ctypes binding and tests/.  See INTEGRATION.md.
// Place as blt_core/build.rs (the reference has no build script: blt_core/Cargo.toml:6-18).
fn main() {
    // libblt_cuda.so is built by `python __graft_entry__.py` (nvcc, sm_100a) in the blt-b200 tree
    let dir = std::env::var("BLT_CUDA_LIB_DIR").expect("set BLT_CUDA_LIB_DIR to blt_b200/lib");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=blt_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=BLT_CUDA_LIB_DIR");
}
