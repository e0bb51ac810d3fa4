"""ctypes binding of libblt_cuda.so (include/blt_cuda.h).  The library is the product: if it is
missing or no CUDA device is usable every call here raises -- there is no Python or CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libblt_cuda.so")

BLT_OK = 0
ERR_NOT_FOUND, ERR_INVALID_INPUT, ERR_INVALID_DATA, ERR_IO = -1, -2, -3, -4
ERR_CUDA, ERR_NOMEM, ERR_CAPACITY, ERR_NO_DEVICE = -5, -6, -7, -8
CONTENT_NONE, CONTENT_TEXT, CONTENT_AUDIO, CONTENT_BIN, CONTENT_VIDEO = -1, 0, 1, 2, 3


class BltError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[blt {code}] {message}")
        self.code = code
        self.message = message


class CoreConfig(C.Structure):
    """blt_core_config == CoreConfig (blt_core/src/lib.rs:110-130) + num_gpus."""
    _fields_ = [("input", C.c_char_p), ("output", C.c_char_p), ("merges_file", C.c_char_p),
                ("content_type", C.c_int), ("has_threads", C.c_int), ("threads", C.c_size_t),
                ("chunk_size", C.c_char_p), ("has_memcap", C.c_int), ("memcap", C.c_uint),
                ("passthrough", C.c_int), ("num_gpus", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                          "blt_b200 has no fallback implementation.")
    L = C.CDLL(LIB_PATH)
    vp, szp, u16p = C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_uint16)
    L.blt_version.restype = C.c_char_p
    L.blt_last_error.restype = C.c_char_p
    L.blt_device_count.argtypes = [C.POINTER(C.c_int)]
    L.blt_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.blt_ctx_destroy.argtypes = [vp]
    L.blt_ctx_destroy.restype = None
    L.blt_strategy_basic.argtypes = [vp, C.POINTER(vp)]
    L.blt_strategy_passthrough.argtypes = [vp, C.POINTER(vp)]
    L.blt_strategy_bpe_from_file.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.blt_strategy_bpe_from_pairs.argtypes = [vp, u16p, u16p, u16p, C.c_size_t, C.POINTER(vp)]
    L.blt_strategy_destroy.argtypes = [vp]
    L.blt_strategy_destroy.restype = None
    L.blt_strategy_num_merges.argtypes = [vp]
    L.blt_strategy_num_merges.restype = C.c_size_t
    L.blt_process_chunk.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, szp]
    L.blt_tokenize_host.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int, vp, C.c_size_t, szp]
    L.blt_process_resident.argtypes = [vp, vp, C.c_size_t, C.c_size_t, vp, C.c_size_t, vp, vp, szp]
    L.blt_resident_result.argtypes = [vp, vp, szp, C.POINTER(C.c_uint32)]
    L.blt_detokenize_host.argtypes = [vp, vp, C.c_size_t, C.c_int, vp, C.c_size_t, szp]
    L.blt_detokenize_resident.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, vp, szp]
    L.blt_count_pairs_host.argtypes = [vp, vp, C.c_size_t, vp]
    L.blt_count_pairs_resident.argtypes = [vp, vp, C.c_size_t, vp, vp]
    L.blt_select_merges.argtypes = [vp, C.c_size_t, C.c_int, vp, vp, szp]
    L.blt_run_tokenizer.argtypes = [C.POINTER(CoreConfig)]
    L.blt_run_detokenizer.argtypes = [C.POINTER(CoreConfig)]
    L.blt_load_bpe_merges.argtypes = [C.c_char_p, u16p, u16p, u16p, C.c_size_t, szp]
    L.blt_parse_chunk_size.argtypes = [C.c_char_p, szp]
    L.blt_effective_chunk_size.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_uint, C.c_uint64]
    L.blt_effective_chunk_size.restype = C.c_size_t
    L.blt_determine_thread_count.argtypes = [C.c_int, C.c_size_t]
    L.blt_determine_thread_count.restype = C.c_size_t
    L.blt_content_type_token.argtypes = [C.c_int]
    L.blt_content_type_token.restype = C.c_uint16
    L.blt_shard_chunks.argtypes = [C.c_size_t, C.c_int, szp]
    L.blt_shard_chunks.restype = None
    L.blt_file_chunk_device.argtypes = [C.c_size_t, C.c_int]
    L.blt_file_chunk_device.restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != BLT_OK:
        raise BltError(rc, lib().blt_last_error().decode("utf-8", "replace"))


def version() -> str:
    return lib().blt_version().decode()


def device_count() -> int:
    n = C.c_int()
    check(lib().blt_device_count(C.byref(n)))
    return n.value


# ---- host-only helpers ---------------------------------------------------------------------------
def load_bpe_merges(path: str) -> Dict[Tuple[int, int], int]:
    n = C.c_size_t()
    check(lib().blt_load_bpe_merges(os.fsencode(path), None, None, None, 0, C.byref(n)))
    k = max(n.value, 1)
    l, r, v = (C.c_uint16 * k)(), (C.c_uint16 * k)(), (C.c_uint16 * k)()
    check(lib().blt_load_bpe_merges(os.fsencode(path), l, r, v, k, C.byref(n)))
    return {(l[i], r[i]): v[i] for i in range(n.value)}


def parse_chunk_size(s: str) -> int:
    out = C.c_size_t()
    check(lib().blt_parse_chunk_size(s.encode(), C.byref(out)))
    return out.value


def effective_chunk_size(cli: Optional[int], threads: int, memcap: int, total_ram: int = 0) -> int:
    return lib().blt_effective_chunk_size(0 if cli is None else 1, cli or 0, threads, memcap, total_ram)


def determine_thread_count(override: Optional[int]) -> int:
    return lib().blt_determine_thread_count(0 if override is None else 1, override or 0)


def content_type_token(ct: int) -> int:
    return lib().blt_content_type_token(ct)


def shard_chunks(n_chunks: int, n_gpus: int):
    b = (C.c_size_t * (n_gpus + 1))()
    lib().blt_shard_chunks(n_chunks, n_gpus, b)
    return list(b)


def file_chunk_device(chunk_index: int, n_gpus: int) -> int:
    """The device blt_run_tokenizer gives chunk `chunk_index` on n_gpus devices (round-robin)."""
    return lib().blt_file_chunk_device(chunk_index, n_gpus)


def run_tokenizer(input: Optional[str], output: Optional[str], merges_file: Optional[str] = None,
                  content_type: int = CONTENT_NONE, threads: Optional[int] = None, chunk_size: Optional[str] = None,
                  memcap: Optional[int] = None, passthrough: bool = False, num_gpus: int = 0) -> None:
    """run_tokenizer(CoreConfig::new_from_cli(...)) (blt_core/src/lib.rs:149-174, 245-267)."""
    cfg = CoreConfig(os.fsencode(input) if input is not None else None,
                     os.fsencode(output) if output is not None else None,
                     os.fsencode(merges_file) if merges_file is not None else None,
                     content_type, 0 if threads is None else 1, threads or 0,
                     chunk_size.encode() if chunk_size is not None else None,
                     0 if memcap is None else 1, memcap or 0, 1 if passthrough else 0, num_gpus)
    check(lib().blt_run_tokenizer(C.byref(cfg)))


def run_detokenizer(input: Optional[str], output: Optional[str], merges_file: Optional[str] = None,
                    content_type: int = CONTENT_NONE, passthrough: bool = False) -> None:
    """File-to-file inverse of run_tokenizer (no reference counterpart)."""
    cfg = CoreConfig(os.fsencode(input) if input is not None else None,
                     os.fsencode(output) if output is not None else None,
                     os.fsencode(merges_file) if merges_file is not None else None,
                     content_type, 0, 0, None, 0, 0, 1 if passthrough else 0, 1)
    check(lib().blt_run_detokenizer(C.byref(cfg)))


def select_merges(counts: np.ndarray, k: int, pad_unobserved: bool = False):
    c = np.ascontiguousarray(counts, dtype=np.uint64)
    left, right = np.empty(max(k, 1), dtype=np.uint8), np.empty(max(k, 1), dtype=np.uint8)
    n = C.c_size_t()
    check(lib().blt_select_merges(c.ctypes.data, k, 1 if pad_unobserved else 0, left.ctypes.data, right.ctypes.data, C.byref(n)))
    return left[: n.value], right[: n.value]


# ---- device objects --------------------------------------------------------------------------------
def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    return np.frombuffer(bytes(data), dtype=np.uint8)


class Strategy:
    """A blt_strategy* == Arc<dyn TokenizationStrategy> (blt_core/src/tokenizer.rs:21-31)."""

    def __init__(self, ctx: "Context", handle: int):
        self.ctx = ctx
        self._h = handle

    @property
    def num_merges(self) -> int:
        return lib().blt_strategy_num_merges(self._h)

    def process_chunk(self, data, out: Optional[np.ndarray] = None) -> np.ndarray:
        """TokenizationStrategy::process_chunk on host buffers."""
        a = _as_u8(data)
        if out is None:
            out = np.empty(max(2 * a.size, 1), dtype=np.uint8)
        n_out = C.c_size_t()
        check(lib().blt_process_chunk(self._h, a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(n_out)))
        return out[: n_out.value]

    def tokenize_host(self, data, chunk_size: int, content_type: int = CONTENT_NONE,
                      out: Optional[np.ndarray] = None) -> np.ndarray:
        """The mmap pipeline on host memory (pipeline.rs:56-131)."""
        a = _as_u8(data)
        if out is None:
            out = np.empty(2 * a.size + 2, dtype=np.uint8)
        n_out = C.c_size_t()
        check(lib().blt_tokenize_host(self._h, a.ctypes.data, a.size, chunk_size, content_type, out.ctypes.data,
                                      out.size, C.byref(n_out)))
        return out[: n_out.value]

    def tokenize_host_ptr(self, in_ptr: int, n: int, chunk_size: int, out_ptr: int, out_cap: int,
                          content_type: int = CONTENT_NONE) -> int:
        n_out = C.c_size_t()
        check(lib().blt_tokenize_host(self._h, in_ptr, n, chunk_size, content_type, out_ptr, out_cap, C.byref(n_out)))
        return n_out.value

    def process_resident(self, d_in: int, n: int, chunk_size: int, d_out: int, out_cap: int,
                         d_chunk_ends: int = 0, stream: int = 0, sync: bool = True) -> Optional[int]:
        """Device pointers in, device pointers out; enqueued on `stream` (a cudaStream_t)."""
        n_out = C.c_size_t()
        check(lib().blt_process_resident(self._h, d_in, n, chunk_size, d_out, out_cap, d_chunk_ends or None,
                                         stream or None, C.byref(n_out) if sync else None))
        return n_out.value if sync else None

    def detokenize_host(self, tokens, has_content_type: bool = False, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Big-endian u16 tokens (as written by tokenize_host / the CLI) back to bytes."""
        a = _as_u8(tokens)
        if out is None:
            out = np.empty(max(a.size, 1), dtype=np.uint8)
        n_out = C.c_size_t()
        check(lib().blt_detokenize_host(self._h, a.ctypes.data, a.size, 1 if has_content_type else 0, out.ctypes.data,
                                        out.size, C.byref(n_out)))
        return out[: n_out.value]

    def detokenize_resident(self, d_tokens: int, n_bytes: int, d_out: int, out_cap: int, stream: int = 0,
                            sync: bool = True) -> Optional[int]:
        n_out = C.c_size_t()
        check(lib().blt_detokenize_resident(self._h, d_tokens, n_bytes, d_out, out_cap, stream or None,
                                            C.byref(n_out) if sync else None))
        return n_out.value if sync else None

    def resident_result(self, stream: int = 0) -> Tuple[int, int]:
        n_out, sweeps = C.c_size_t(), C.c_uint32()
        check(lib().blt_resident_result(self._h, stream or None, C.byref(n_out), C.byref(sweeps)))
        return n_out.value, sweeps.value

    def close(self) -> None:
        if self._h:
            lib().blt_strategy_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """A blt_ctx*: one CUDA device."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        check(lib().blt_ctx_create(device, C.byref(h)))
        self._h = h.value
        self.device = device

    def _wrap(self, rc: int, h: C.c_void_p) -> Strategy:
        check(rc)
        return Strategy(self, h.value)

    def basic(self) -> Strategy:
        h = C.c_void_p()
        return self._wrap(lib().blt_strategy_basic(self._h, C.byref(h)), h)

    def passthrough(self) -> Strategy:
        h = C.c_void_p()
        return self._wrap(lib().blt_strategy_passthrough(self._h, C.byref(h)), h)

    def bpe_from_file(self, path: str) -> Strategy:
        h = C.c_void_p()
        return self._wrap(lib().blt_strategy_bpe_from_file(self._h, os.fsencode(path), C.byref(h)), h)

    def bpe_from_pairs(self, pairs: Dict[Tuple[int, int], int]) -> Strategy:
        n = len(pairs)
        k = max(n, 1)
        l, r, v = (C.c_uint16 * k)(), (C.c_uint16 * k)(), (C.c_uint16 * k)()
        for i, ((a, b), val) in enumerate(pairs.items()):
            l[i], r[i], v[i] = a, b, val
        h = C.c_void_p()
        return self._wrap(lib().blt_strategy_bpe_from_pairs(self._h, l, r, v, n, C.byref(h)), h)

    def count_pairs(self, data) -> np.ndarray:
        """Adjacent-byte-pair histogram on the GPU: counts[b0 << 8 | b1] (uint64, 65 536 entries)."""
        a = _as_u8(data)
        counts = np.zeros(65536, dtype=np.uint64)
        check(lib().blt_count_pairs_host(self._h, a.ctypes.data, a.size, counts.ctypes.data))
        return counts

    def count_pairs_resident(self, d_in: int, n: int, d_counts: int, stream: int = 0) -> None:
        check(lib().blt_count_pairs_resident(self._h, d_in, n, d_counts, stream or None))

    def train_merges(self, data, k: int, pad_unobserved: bool = False):
        """The k most frequent adjacent byte pairs of `data`, in merges.txt order (ids 256, 257, ...)."""
        return select_merges(self.count_pairs(data), k, pad_unobserved)

    def close(self) -> None:
        if self._h:
            lib().blt_ctx_destroy(self._h)
            self._h = None
