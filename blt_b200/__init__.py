"""blt_b200 -- B200 (sm_100a) implementation of blt's tokenization hot path, with the Python surface
of the reference's `blt` module (blt_python/src/lib.rs:27-220, blt_python/python/blt/__init__.py):
ByteTokenizer, load_bpe_merges, version, __version__.  All work goes through libblt_cuda.so; if the
library or a CUDA device is missing, calls raise (no fallback)."""
from __future__ import annotations

import os
import tempfile
from typing import Dict, Optional, Tuple

from . import _native

__all__ = ["ByteTokenizer", "load_bpe_merges", "version", "__version__"]


def version() -> str:
    """blt.version() (blt_python/src/lib.rs:205-208)."""
    return _native.version()


def load_bpe_merges(path: str) -> Dict[Tuple[int, int], int]:
    """blt.load_bpe_merges(path) (blt_python/src/lib.rs:194-198, blt_core/src/lib.rs:216-230).
    IOError (OSError) if the file cannot be read, ValueError if a line is malformed."""
    try:
        return _native.load_bpe_merges(path)
    except _native.BltError as e:
        if e.code == _native.ERR_NOT_FOUND:
            raise FileNotFoundError(e.message) from None
        if e.code == _native.ERR_INVALID_DATA:
            raise ValueError(e.message) from None
        raise OSError(e.message) from None


def _rust_opt(v) -> str:
    """Rust `{:?}` of an Option<T> for T in {String, usize, u8}."""
    if v is None:
        return "None"
    if isinstance(v, str):
        return 'Some("%s")' % v.replace("\\", "\\\\").replace('"', '\\"')
    return f"Some({v})"


class ByteTokenizer:
    """blt.ByteTokenizer (blt_python/src/lib.rs:27-178).

    merges: optional {(byte1, byte2): id}.  As in the reference, only the KEYS are used: they are
    written to a temporary merges file and ids are assigned 256, 257, ... in file order
    (blt_python/src/lib.rs:104-114).  The reference writes them in HashMap iteration order, which
    is randomised per process; this build writes them in the dict's insertion order, which makes
    the ids deterministic."""

    def __init__(self, merges: Optional[Dict[Tuple[int, int], int]] = None, content_type: Optional[str] = None,
                 threads: Optional[int] = None, chunk_size: Optional[str] = None, memory_cap: Optional[int] = None):
        if memory_cap is not None:
            if not (0 <= int(memory_cap) <= 255):
                raise OverflowError("memory_cap out of range for u8")       # PyO3 u8 extraction
            if memory_cap > 100:
                raise ValueError("memory_cap must be between 0 and 100")    # lib.rs:57-64
        if content_type is not None and content_type not in ("Text", "Bin"):
            raise ValueError("content_type must be 'Text' or 'Bin'")        # lib.rs:66-75
        if threads is not None and int(threads) < 0:
            raise OverflowError("can't convert negative int to unsigned")   # PyO3 usize extraction
        if merges is not None:
            for k, v in merges.items():
                a, b = k
                if not (0 <= int(a) <= 255 and 0 <= int(b) <= 255 and 0 <= int(v) <= 65535):
                    raise OverflowError("merges must map (u8, u8) -> u16")
        self.merges = dict(merges) if merges is not None else None
        self.content_type = content_type
        self.threads = threads
        self.chunk_size = chunk_size
        self.memory_cap = memory_cap

    def tokenize_file(self, input_path: str, output_path: str) -> None:
        """ByteTokenizer.tokenize_file (blt_python/src/lib.rs:98-165)."""
        self._run(input_path, output_path, inverse=False)

    def detokenize_file(self, input_path: str, output_path: str) -> None:
        """The inverse of tokenize_file with the same merges / content_type (an addition: the reference has
        no detokenizer): input_path holds the tokens, output_path receives the bytes."""
        self._run(input_path, output_path, inverse=True)

    def _run(self, input_path: str, output_path: str, inverse: bool) -> None:
        ct = {None: _native.CONTENT_NONE, "Text": _native.CONTENT_TEXT, "Bin": _native.CONTENT_BIN}[self.content_type]
        tmp = None
        try:
            if self.merges is not None:
                fd, tmp = tempfile.mkstemp(prefix="blt_merges_", suffix=".txt")
                with os.fdopen(fd, "w") as f:
                    f.write("".join(f"{a} {b}\n" for (a, b) in self.merges.keys()))
            try:
                if inverse:
                    _native.run_detokenizer(input_path, output_path, tmp, ct)
                else:
                    _native.run_tokenizer(input_path, output_path, tmp, ct, self.threads, self.chunk_size,
                                          self.memory_cap, passthrough=False)
            except _native.BltError as e:
                if e.code == _native.ERR_NOT_FOUND:
                    raise FileNotFoundError(e.message) from None
                if e.code in (_native.ERR_INVALID_INPUT, _native.ERR_INVALID_DATA):
                    raise ValueError(e.message) from None
                if e.code == _native.ERR_IO:
                    raise OSError(e.message) from None
                raise RuntimeError(e.message) from None
        finally:
            if tmp is not None:
                os.unlink(tmp)

    def __repr__(self) -> str:  # lib.rs:168-177
        return ("ByteTokenizer(merges=%d, content_type=%s, threads=%s, chunk_size=%s, memory_cap=%s)"
                % (len(self.merges) if self.merges is not None else 0, _rust_opt(self.content_type),
                   _rust_opt(self.threads), _rust_opt(self.chunk_size), _rust_opt(self.memory_cap)))


def __getattr__(name):
    # `blt.__version__` (blt_python/python/blt/__init__.py:16), resolved on first use: importing the package (for
    # instance for blt_b200.synth, the workload generators) must not load libblt_cuda.so.
    if name == "__version__":
        return version()
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
