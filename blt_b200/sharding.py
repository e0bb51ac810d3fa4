"""Multi-process form of the chunk pipeline: one process per GPU, chunks sharded contiguously
(`chunk k -> rank floor(k*W/K)`), no data-path collective.  The only thing ranks
exchange is each shard's output length (an all_gather of one integer), from which every rank derives
the file offset of its shard; the token bytes themselves never leave the rank that produced them."""
from __future__ import annotations

import mmap
import os
from typing import Callable, List, Optional, Tuple

from . import _native


def shard_range(n_chunks: int, world: int, rank: int) -> Tuple[int, int]:
    b = _native.shard_chunks(n_chunks, world)
    return b[rank], b[rank + 1]


def output_offsets(lengths: List[int], prefix: int = 0) -> List[int]:
    """Exclusive prefix of the per-rank output lengths (plus the content-type prefix)."""
    out, acc = [], prefix
    for n in lengths:
        out.append(acc)
        acc += n
    return out


def tokenize_file_sharded(process_chunk: Callable[[memoryview], bytes], in_path: str, out_path: str, chunk_size: int,
                          rank: int, world: int, content_type_token: Optional[int] = None, dist=None) -> int:
    """Every rank tokenizes chunks [first,last) of the input with `process_chunk` (on a GPU box:
    Strategy.process_chunk; in CPU tests: a stand-in), learns its file offset from an all_gather of
    the shard lengths and writes its shard with pwrite.  Returns the total output length."""
    size = os.path.getsize(in_path)
    n_chunks = (size + chunk_size - 1) // chunk_size if size else 0
    first, last = shard_range(n_chunks, world, rank) if n_chunks else (0, 0)
    parts: List[bytes] = []
    if last > first:
        with open(in_path, "rb") as f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            view = memoryview(mm)
            for k in range(first, last):
                parts.append(bytes(process_chunk(view[k * chunk_size:min((k + 1) * chunk_size, size)])))
            view.release()
    local = sum(len(p) for p in parts)
    if dist is not None and world > 1:
        import torch
        t = torch.tensor([local], dtype=torch.int64)
        gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, t)
        lengths = [int(g.item()) for g in gathered]
    else:
        lengths = [local]
    prefix = 2 if content_type_token is not None else 0
    offsets = output_offsets(lengths, prefix)
    if rank == 0:
        with open(out_path, "wb") as f:  # File::create truncates (io_handler.rs:70)
            if content_type_token is not None:
                f.write(int(content_type_token).to_bytes(2, "big"))
    if dist is not None and world > 1:
        dist.barrier()
    fd = os.open(out_path, os.O_WRONLY)
    try:
        off = offsets[rank]
        for p in parts:
            os.pwrite(fd, p, off)
            off += len(p)
    finally:
        os.close(fd)
    if dist is not None and world > 1:
        dist.barrier()
    return prefix + sum(lengths)
