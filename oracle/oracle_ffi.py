"""ctypes access to the C++ CPU oracle (oracle/blt_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs and
by nothing else.  PARITY PINNING: see blt_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libblt_oracle.so")

ORA_OK, ORA_NOT_FOUND, ORA_INVALID_INPUT, ORA_INVALID_DATA, ORA_IO, ORA_CAPACITY = 0, -1, -2, -3, -4, -7
MODE_BASIC, MODE_BPE, MODE_PASSTHROUGH = 0, 1, 2
_MODES = {"basic": MODE_BASIC, "bpe": MODE_BPE, "passthrough": MODE_PASSTHROUGH}


class OracleError(Exception):
    def __init__(self, kind: int, message: str):
        super().__init__(f"[{kind}] {message}")
        self.kind = kind
        self.message = message


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, no reference sources needed)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("blt_oracle.cpp", "blt_oracle.h", "Makefile"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_m:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    u8p, u16p, szp = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_size_t)
    L.ora_merges_new.restype = C.c_void_p
    L.ora_merges_free.argtypes = [C.c_void_p]
    L.ora_merges_insert.argtypes = [C.c_void_p, C.c_uint16, C.c_uint16, C.c_uint16]
    L.ora_merges_len.argtypes = [C.c_void_p]
    L.ora_merges_len.restype = C.c_size_t
    L.ora_merges_export.argtypes = [C.c_void_p, u16p, u16p, u16p, C.c_size_t]
    L.ora_merges_export.restype = C.c_size_t
    L.ora_load_bpe_merges.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    L.ora_bpe_process_chunk.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, szp,
                                        C.POINTER(C.c_uint32)]
    L.ora_basic_process_chunk.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, szp]
    L.ora_passthrough_process_chunk.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, szp]
    L.ora_run_buffer.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                 C.c_void_p, C.c_size_t, szp]
    L.ora_run_files.argtypes = [C.c_int, C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_size_t, C.c_int,
                                C.c_char_p, C.c_size_t]
    L.ora_parse_chunk_size.argtypes = [C.c_char_p, szp, C.c_char_p, C.c_size_t]
    L.ora_effective_chunk_size.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_uint, C.c_uint64]
    L.ora_effective_chunk_size.restype = C.c_size_t
    L.ora_determine_thread_count.argtypes = [C.c_int, C.c_size_t, C.c_size_t]
    L.ora_determine_thread_count.restype = C.c_size_t
    L.ora_content_type_token.argtypes = [C.c_int]
    L.ora_content_type_token.restype = C.c_uint16
    L.ora_detokenize.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, szp]
    _lib = L
    return L


class Merges:
    """Owns an ora_merges* (BpeMerges, lib.rs:75)."""

    def __init__(self, pairs: Optional[Dict[Tuple[int, int], int]] = None, _handle=None):
        self._h = _handle if _handle is not None else lib().ora_merges_new()
        if pairs:
            for (a, b), v in pairs.items():
                lib().ora_merges_insert(self._h, a, b, v)

    @classmethod
    def from_file(cls, path: str) -> "Merges":
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = lib().ora_load_bpe_merges(os.fsencode(path), C.byref(h), err, len(err))
        if rc != ORA_OK:
            raise OracleError(rc, err.value.decode("utf-8", "replace"))
        return cls(_handle=h.value)

    def __len__(self) -> int:
        return lib().ora_merges_len(self._h)

    def to_dict(self) -> Dict[Tuple[int, int], int]:
        n = len(self)
        l, r, v = (C.c_uint16 * max(n, 1))(), (C.c_uint16 * max(n, 1))(), (C.c_uint16 * max(n, 1))()
        k = lib().ora_merges_export(self._h, l, r, v, n)
        return {(l[i], r[i]): v[i] for i in range(k)}

    def __del__(self):
        try:
            if self._h:
                lib().ora_merges_free(self._h)
                self._h = None
        except Exception:
            pass


def _buf(data) -> Tuple[C.c_void_p, int, object]:
    """Accept bytes / bytearray / numpy uint8 array without copying where possible."""
    try:
        import numpy as np
        if isinstance(data, np.ndarray):
            a = np.ascontiguousarray(data, dtype=np.uint8)
            return C.c_void_p(a.ctypes.data), a.size, a
    except ImportError:  # pragma: no cover
        pass
    b = bytes(data)
    return C.cast(C.c_char_p(b), C.c_void_p), len(b), b


def process_chunk(mode: str, data, merges: Optional[Merges] = None, want_sweeps: bool = False):
    """One strategy call (tokenizer.rs process_chunk).  Returns bytes (and sweeps for BPE)."""
    p, n, keep = _buf(data)
    cap = max(2 * n, 1)
    out = C.create_string_buffer(cap)
    out_len = C.c_size_t()
    sweeps = C.c_uint32()
    if mode == "bpe":
        rc = lib().ora_bpe_process_chunk(merges._h, p, n, out, cap, C.byref(out_len), C.byref(sweeps))
    elif mode == "basic":
        rc = lib().ora_basic_process_chunk(p, n, out, cap, C.byref(out_len))
    else:
        rc = lib().ora_passthrough_process_chunk(p, n, out, cap, C.byref(out_len))
    if rc != ORA_OK:
        raise OracleError(rc, "process_chunk failed")
    res = out.raw[: out_len.value]
    return (res, sweeps.value) if want_sweeps else res


def run_buffer(mode: str, data, chunk_size: int, threads: int = 1, merges: Optional[Merges] = None,
               content_type_token: Optional[int] = None, out_array=None):
    """In-memory run_tokenizer.  Returns a numpy uint8 array view of the output."""
    import numpy as np
    p, n, keep = _buf(data)
    cap = 2 * n + 2
    out = out_array if out_array is not None else np.empty(max(cap, 1), dtype=np.uint8)
    assert out.size >= cap or n == 0
    out_len = C.c_size_t()
    rc = lib().ora_run_buffer(_MODES[mode], merges._h if merges is not None else None, p, n, chunk_size, threads,
                              -1 if content_type_token is None else content_type_token,
                              C.c_void_p(out.ctypes.data), out.size, C.byref(out_len))
    if rc != ORA_OK:
        raise OracleError(rc, "run_buffer failed")
    return out[: out_len.value]


def detokenize(tokens, merges: Optional[Merges] = None, has_content_type: bool = False):
    """Inverse of the wire format (no reference counterpart, see blt_oracle.h).  Returns a numpy uint8 array."""
    import numpy as np
    p, n, keep = _buf(tokens)
    out = np.empty(max(n, 1), dtype=np.uint8)
    out_len = C.c_size_t()
    rc = lib().ora_detokenize(merges._h if merges is not None else None, p, n, 1 if has_content_type else 0,
                              C.c_void_p(out.ctypes.data), out.size, C.byref(out_len))
    if rc != ORA_OK:
        raise OracleError(rc, "detokenize failed")
    return out[: out_len.value]


def run_files(mode: str, in_path: str, out_path: str, chunk_size: int, threads: int = 1,
              merges: Optional[Merges] = None, content_type_token: Optional[int] = None) -> None:
    err = C.create_string_buffer(1024)
    rc = lib().ora_run_files(_MODES[mode], merges._h if merges is not None else None, os.fsencode(in_path),
                             os.fsencode(out_path), chunk_size, threads,
                             -1 if content_type_token is None else content_type_token, err, len(err))
    if rc != ORA_OK:
        raise OracleError(rc, err.value.decode("utf-8", "replace"))


def parse_chunk_size(s: str) -> int:
    out = C.c_size_t()
    err = C.create_string_buffer(512)
    rc = lib().ora_parse_chunk_size(s.encode(), C.byref(out), err, len(err))
    if rc != ORA_OK:
        raise OracleError(rc, err.value.decode())
    return out.value


def effective_chunk_size(cli: Optional[int], threads: int, memcap: int, total_ram: int) -> int:
    return lib().ora_effective_chunk_size(0 if cli is None else 1, cli or 0, threads, memcap, total_ram)


def determine_thread_count(override: Optional[int], logical_cpus: int) -> int:
    return lib().ora_determine_thread_count(0 if override is None else 1, override or 0, logical_cpus)


def content_type_token(ct: int) -> int:
    return lib().ora_content_type_token(ct)
