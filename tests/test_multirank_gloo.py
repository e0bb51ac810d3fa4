"""World-size-2 gloo test (CPU) of the N>1 path: contiguous chunk sharding, the all_gather of shard
lengths, the offset prefix and the pwrite assembly of blt_b200/sharding.py, plus the timing reduction
bench.py uses (max over ranks).  The per-chunk compute is the CPU oracle standing in for the GPU
strategy (tests may use the oracle; the product path never does)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, in_path, out_path, merges, chunk, result_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from blt_b200 import sharding
    from oracle import oracle_ffi as ora
    m = ora.Merges(merges)
    total = sharding.tokenize_file_sharded(lambda v: ora.process_chunk("bpe", np.frombuffer(v, dtype=np.uint8), m),
                                           in_path, out_path, chunk, rank, world, 0xFF01, dist)
    # bench.py's timing rule: the job time is the max over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        with open(result_path, "w") as f:
            f.write(f"{total} {t.item()}")
    dist.destroy_process_group()


@pytest.mark.parametrize("n,chunk", [(1_000_003, 65536), (5 * 65536, 65536), (100, 65536), (0, 4096)])
def test_two_ranks_assemble_the_reference_output(tmp_path, oracle, n, chunk):
    rng = np.random.default_rng(n + 1)
    data = rng.choice(np.frombuffer(b"ab c", dtype=np.uint8), size=n)
    merges = {(97, 98): 256, (98, 97): 257, (97, 97): 258, (32, 97): 259, (99, 32): 260}
    (tmp_path / "in.bin").write_bytes(data.tobytes())
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), merges, chunk,
                            str(tmp_path / "res.txt")), nprocs=2, join=True)
    want = bytes(oracle.run_buffer("bpe", data, chunk, 2, oracle.Merges(merges), 0xFF01))
    assert (tmp_path / "out.bin").read_bytes() == want
    total, tmax = (tmp_path / "res.txt").read_text().split()
    assert int(total) == len(want) and float(tmax) == 11.0


def test_shard_ranges_cover_all_chunks():
    from blt_b200 import sharding
    for n_chunks in (0, 1, 2, 7, 64, 513):
        for world in (1, 2, 4, 8):
            if n_chunks == 0:
                continue
            spans = [sharding.shard_range(n_chunks, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_chunks
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert sharding.output_offsets([5, 0, 7], 2) == [2, 7, 7]
