// UNTESTED: there is no Rust toolchain in the image this repository is built and tested in (no rustc, cargo or
// maturin).  Everything this file calls is exercised through the same C ABI (include/blt_cuda.h) by the C++ CLI,
// the ctypes binding and tests/.  See INTEGRATION.md.
// Place as blt_core/src/cuda_strategy.rs and add `mod cuda_strategy;` to blt_core/src/lib.rs.
// Implements the reference's operator interface (blt_core/src/tokenizer.rs:21-31) on top of blt_process_chunk.
use crate::tokenizer::TokenizationStrategy;
use crate::BpeMerges;
use std::{ffi::{c_char, c_int, CStr, CString}, io, path::Path, ptr};

#[repr(C)] pub struct blt_ctx { _p: [u8; 0] }
#[repr(C)] pub struct blt_strategy { _p: [u8; 0] }

extern "C" {
    fn blt_last_error() -> *const c_char;
    fn blt_ctx_create(device: c_int, out: *mut *mut blt_ctx) -> c_int;
    fn blt_ctx_destroy(ctx: *mut blt_ctx);
    fn blt_strategy_basic(ctx: *mut blt_ctx, out: *mut *mut blt_strategy) -> c_int;
    fn blt_strategy_passthrough(ctx: *mut blt_ctx, out: *mut *mut blt_strategy) -> c_int;
    fn blt_strategy_bpe_from_file(ctx: *mut blt_ctx, path: *const c_char, out: *mut *mut blt_strategy) -> c_int;
    fn blt_strategy_bpe_from_pairs(ctx: *mut blt_ctx, l: *const u16, r: *const u16, v: *const u16, n: usize,
                                   out: *mut *mut blt_strategy) -> c_int;
    fn blt_strategy_destroy(s: *mut blt_strategy);
    fn blt_process_chunk(s: *mut blt_strategy, input: *const u8, n: usize, out: *mut u8, out_cap: usize,
                         out_len: *mut usize) -> c_int;
}

fn to_io(code: c_int) -> io::Error {
    let msg = unsafe { CStr::from_ptr(blt_last_error()) }.to_string_lossy().into_owned();
    let kind = match code { -1 => io::ErrorKind::NotFound, -2 => io::ErrorKind::InvalidInput,
                            -3 => io::ErrorKind::InvalidData, _ => io::ErrorKind::Other };
    io::Error::new(kind, msg)
}

/// Owns one device context and one strategy handle.  `blt_process_chunk` is re-entrant on a
/// handle (it leases a private stream/buffer set per call), which is what `Send + Sync` needs.
pub struct CudaStrategy { ctx: *mut blt_ctx, s: *mut blt_strategy }
unsafe impl Send for CudaStrategy {}
unsafe impl Sync for CudaStrategy {}

impl CudaStrategy {
    fn with(device: i32, make: impl FnOnce(*mut blt_ctx, *mut *mut blt_strategy) -> c_int) -> io::Result<Self> {
        let (mut ctx, mut s) = (ptr::null_mut(), ptr::null_mut());
        let rc = unsafe { blt_ctx_create(device, &mut ctx) };
        if rc != 0 { return Err(to_io(rc)); }
        let rc = make(ctx, &mut s);
        if rc != 0 { unsafe { blt_ctx_destroy(ctx) }; return Err(to_io(rc)); }
        Ok(Self { ctx, s })
    }
    pub fn basic(device: i32) -> io::Result<Self> { Self::with(device, |c, o| unsafe { blt_strategy_basic(c, o) }) }
    pub fn passthrough(device: i32) -> io::Result<Self> { Self::with(device, |c, o| unsafe { blt_strategy_passthrough(c, o) }) }
    pub fn bpe_from_file(device: i32, path: &Path) -> io::Result<Self> {
        let p = CString::new(path.to_string_lossy().as_bytes()).map_err(|e| io::Error::new(io::ErrorKind::InvalidInput, e))?;
        Self::with(device, |c, o| unsafe { blt_strategy_bpe_from_file(c, p.as_ptr(), o) })
    }
    pub fn bpe(device: i32, merges: &BpeMerges) -> io::Result<Self> {
        let (mut l, mut r, mut v) = (Vec::new(), Vec::new(), Vec::new());
        for (&(a, b), &id) in merges { l.push(a); r.push(b); v.push(id); }
        Self::with(device, |c, o| unsafe { blt_strategy_bpe_from_pairs(c, l.as_ptr(), r.as_ptr(), v.as_ptr(), l.len(), o) })
    }
}

impl Drop for CudaStrategy {
    fn drop(&mut self) { unsafe { blt_strategy_destroy(self.s); blt_ctx_destroy(self.ctx); } }
}

#[async_trait::async_trait]
impl TokenizationStrategy for CudaStrategy {
    async fn process_chunk(&self, chunk_data: &[u8]) -> io::Result<Vec<u8>> {
        let mut out = vec![0u8; chunk_data.len() * 2];          // 2*n always suffices
        let mut len = 0usize;
        let rc = unsafe { blt_process_chunk(self.s, chunk_data.as_ptr(), chunk_data.len(), out.as_mut_ptr(), out.len(), &mut len) };
        if rc != 0 { return Err(to_io(rc)); }
        out.truncate(len);
        Ok(out)
    }
}
