"""Independent second restatement of the reference's tokenization semantics, in plain Python.

TEST INFRASTRUCTURE ONLY (same rule as blt_oracle.h): used by tests/ to cross-check the C++ oracle
and the closed parallel form the CUDA kernels implement.  Never imported by the product.

PARITY PINNING: the Rust reference cannot be built here; both this model and the C++ oracle are
checked against every golden vector in the reference's own tests (tests/test_oracle_golden.py).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

Merges = Dict[Tuple[int, int], int]


def bpe_sweep(tokens: List[int], merges: Merges) -> Tuple[List[int], bool]:
    """One pass of the loop body at blt_core/src/tokenizer.rs:64-81."""
    out: List[int] = []
    found = False
    i, n = 0, len(tokens)
    while i < n:
        if i < n - 1 and (tokens[i], tokens[i + 1]) in merges:
            out.append(merges[(tokens[i], tokens[i + 1])])
            i += 2
            found = True
        else:
            out.append(tokens[i])
            i += 1
    return out, found


def bpe_tokens(data: bytes, merges: Merges) -> List[int]:
    """BpeStrategy::process_chunk up to (not including) serialisation, tokenizer.rs:56-86."""
    if not data:
        return []
    tokens = list(data)
    while True:
        tokens, found = bpe_sweep(tokens, merges)
        if not found:
            return tokens


def to_be(tokens: Iterable[int]) -> bytes:
    """u16::to_be_bytes per token, tokenizer.rs:88-91."""
    return b"".join(int(t).to_bytes(2, "big") for t in tokens)


def bpe_chunk(data: bytes, merges: Merges) -> bytes:
    return to_be(bpe_tokens(data, merges))


def basic_chunk(data: bytes) -> bytes:
    """BasicTokenizationStrategy::process_chunk, tokenizer.rs:108-123."""
    return to_be(data)


def passthrough_chunk(data: bytes) -> bytes:
    """PassthroughStrategy::process_chunk, tokenizer.rs:138-144."""
    return bytes(data)


def run_buffer(mode: str, data: bytes, chunk_size: int, merges: Optional[Merges] = None,
               content_type_token: Optional[int] = None) -> bytes:
    """run_tokenizer on an in-memory file: prefix (lib.rs:284-294), fixed-offset chunks
    (pipeline.rs:73-81), concatenation in chunk order (pipeline.rs:153-168)."""
    out = bytearray()
    if content_type_token is not None:
        out += int(content_type_token).to_bytes(2, "big")
    for start in range(0, len(data), chunk_size):
        chunk = data[start:start + chunk_size]
        if mode == "passthrough":
            out += passthrough_chunk(chunk)
        elif mode == "bpe":
            out += bpe_chunk(chunk, merges or {})
        else:
            out += basic_chunk(chunk)
    return bytes(out)


# ---- the closed parallel form the CUDA sweep kernel implements -----

def sweep_parallel_form(tokens: List[int], merges: Merges, chunk_ends: Iterable[int] = ()) -> List[int]:
    """m[i] = pair (t[i],t[i+1]) in map (0 at the last token and at every chunk-last index);
    start[i] = m[i] & ~start[i-1]; emit map[pair] at starts, drop the token after a start, copy
    the rest.  Must equal bpe_sweep() on every input; that equivalence is property-tested."""
    n = len(tokens)
    ends = set(chunk_ends)
    m = [0] * n
    for i in range(n - 1):
        if i not in ends and (tokens[i], tokens[i + 1]) in merges:
            m[i] = 1
    out: List[int] = []
    prev_start = 0
    for i in range(n):
        start = m[i] & (1 - prev_start)
        if start:
            out.append(merges[(tokens[i], tokens[i + 1])])
        elif not prev_start:
            out.append(tokens[i])
        prev_start = start
    return out


def load_bpe_merges_text(text: str) -> Merges:
    """config_loader.rs:14-46 for already-decoded, valid text (error paths live in the C++ oracle)."""
    merges: Merges = {}
    vocab = 256
    for line in text.split("\n"):
        if line.endswith("\r"):
            line = line[:-1]
        if line.startswith("#") or line == "":
            continue
        parts = line.split()
        if len(parts) != 2:
            raise ValueError("Invalid merge rule format")
        a, b = int(parts[0]), int(parts[1])
        if not (0 <= a <= 255 and 0 <= b <= 255):
            raise ValueError("byte out of range")
        merges[(a, b)] = vocab
        vocab += 1
    return merges


def detokenize(stream: bytes, merges: Optional[Merges] = None, has_content_type: bool = False) -> bytes:
    """Inverse of to_be(): token < 256 -> that byte, id of (l, r) -> bytes l r (no reference counterpart;
    second restatement of oracle/blt_oracle.cpp::ora_detokenize)."""
    inverse = {}
    for (l, r), v in (merges or {}).items():
        if l > 255 or r > 255 or v < 256 or v in inverse:
            raise ValueError("table is not invertible")
        inverse[v] = (l, r)
    if len(stream) % 2:
        raise ValueError("odd number of bytes")
    toks = [int.from_bytes(stream[i:i + 2], "big") for i in range(0, len(stream), 2)]
    if has_content_type:
        if not toks or not 0xFF01 <= toks[0] <= 0xFF04:
            raise ValueError("missing content-type token")
        toks = toks[1:]
    out = bytearray()
    for t in toks:
        if t < 256:
            out.append(t)
        else:
            out.extend(inverse[t])   # KeyError: token not in the table
    return bytes(out)
