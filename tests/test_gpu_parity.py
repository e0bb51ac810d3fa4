"""Parity tests proper: the CUDA path, called through the C ABI (libblt_cuda.so), against the CPU
oracle on the same inputs.  Bit-exact everywhere (byte / integer work, no tolerance).  Needs a B200:
run with  python -m pytest tests -m gpu."""
import json
import os
import random
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))
DER = json.load(open(os.path.join(HERE, "golden", "derived_vectors.json")))
BLT = os.path.join(ROOT, "blt_b200", "lib", "blt")
MiB = 1 << 20


def be(tokens):
    return b"".join(int(t).to_bytes(2, "big") for t in tokens)


def merges_dict(rows):
    return {(a, b): v for a, b, v in rows}


@pytest.fixture(scope="module")
def nat():
    from blt_b200 import _native
    return _native


@pytest.fixture(scope="module")
def ctx(nat):
    c = nat.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


def resident(torch, strat, data: np.ndarray, chunk: int, want_ends: bool = True):
    """blt_process_resident on torch-owned device memory; returns (output bytes, chunk_ends)."""
    n = data.size
    d_in = torch.from_numpy(np.ascontiguousarray(data)).cuda() if n else torch.empty(16, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(2 * n + 16, dtype=torch.uint8, device="cuda")
    c = chunk if chunk else max(n, 1)
    n_chunks = max(1, (n + c - 1) // c)
    d_ends = torch.full((n_chunks,), -1, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    out_len = strat.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), 2 * n, d_ends.data_ptr() if want_ends else 0,
                                     stream, sync=True)
    return d_out[:out_len].cpu().numpy(), d_ends.cpu().numpy()


# ---- reference-held golden vectors through the C ABI -------------------------------------------------

@pytest.mark.parametrize("row", REF["bpe"], ids=lambda r: r["src"])
def test_ref_bpe_vectors(ctx, row, tmp_path):
    data = row["input"].encode()
    if "merges_file" in row:
        p = tmp_path / "merges.txt"
        p.write_text(row["merges_file"])
        s = ctx.bpe_from_file(str(p))
    else:
        s = ctx.bpe_from_pairs(merges_dict(row["merges"]))
    assert s.process_chunk(data).tobytes() == be(row["tokens"])
    assert s.tokenize_host(data, chunk_size=1 << 20).tobytes() == be(row["tokens"])


@pytest.mark.parametrize("row", REF["basic"], ids=lambda r: r["src"])
def test_ref_basic_vectors(ctx, row):
    s = ctx.basic()
    assert s.process_chunk(row["input"].encode()).tobytes() == bytes.fromhex(row["bytes_hex"])


def test_ref_passthrough_and_content_type(ctx, nat):
    for row in REF["passthrough"]:
        assert ctx.passthrough().process_chunk(row["input"].encode()).tobytes() == bytes.fromhex(row["bytes_hex"])
    row = REF["content_type"][0]
    got = ctx.basic().tokenize_host(row["input"].encode(), 1 << 20, nat.CONTENT_TEXT)
    assert got.tobytes() == bytes.fromhex(row["bytes_hex"])


# ---- derived vectors: runs, overlaps, chains, cycles, chunk walls -------------------------------------

@pytest.mark.parametrize("block", ["hand", "random"])
def test_derived_vectors(ctx, torch_mod, block):
    for row in DER[block]:
        md = merges_dict(row["merges"])
        data = np.frombuffer(bytes.fromhex(row["input_hex"]), dtype=np.uint8)
        chunk = row["chunk"] or max(data.size, 1)
        s = ctx.bpe_from_pairs(md)
        assert s.tokenize_host(data, chunk_size=chunk).tobytes() == be(row["tokens"]), row
        got, _ = resident(torch_mod, s, data, chunk)
        assert got.tobytes() == be(row["tokens"]), row
        s.close()


# ---- random inputs vs the oracle, every tile configuration ---------------------------------------------

def _random_case(rng, n, alphabet, density):
    data = np.frombuffer(bytes(rng.choice(alphabet) for _ in range(n)), dtype=np.uint8) if n < 4096 else \
        np.random.default_rng(rng.randrange(1 << 30)).choice(np.frombuffer(bytes(alphabet), dtype=np.uint8), size=n)
    pairs = {}
    for a in alphabet:
        for b in alphabet:
            if rng.random() < density:
                pairs[(a, b)] = 256 + len(pairs)
    return np.ascontiguousarray(data), pairs


@pytest.mark.parametrize("variant,dense", [(0, "always"), (1, "always"), (2, "always"), (0, "0"), (1, "0"), (2, "0"), (0, "1")])
def test_random_vs_oracle_all_variants(nat, torch_mod, oracle, variant, dense, monkeypatch):
    """Both tile sizes of the exact sweep; the dense speculative pass attempted on every call (its failure
    launches the exact sweep from the device), never, or as the predictor decides (the default)."""
    monkeypatch.setenv("BLT_SWEEP_VARIANT", str(variant))
    monkeypatch.setenv("BLT_DENSE", dense)
    c = nat.Context(0)
    rng = random.Random(1000 + variant)
    sizes = [1, 2, 15, 16, 17, 255, 4095, 4096, 4097, 8191, 8193, 16384, 65537, 300001, 1 * MiB + 3, 5 * MiB + 11]
    for n in sizes:
        for density in (1.0, 0.85, 0.3):
            alphabet = [97, 98, 99][: rng.choice([1, 2, 3])] if density == 1.0 else [97, 98, 99, 100, 32]
            data, pairs = _random_case(rng, n, alphabet, density)
            om = oracle.Merges(pairs)
            s = c.bpe_from_pairs(pairs)
            for chunk in (0, 4096, 65536, 100000, 1 * MiB):
                if chunk > n and chunk != 0:
                    continue
                want = oracle.run_buffer("bpe", data, chunk or max(n, 1), 4, om)
                got, ends = resident(torch_mod, s, data, chunk)
                assert np.array_equal(got, want), (variant, n, density, chunk)
                # chunk_ends is the inclusive prefix of per-chunk output lengths
                cc = chunk or n
                acc = 0
                if (n + cc - 1) // cc > 64:
                    assert int(ends[-1]) == want.size
                    continue
                for k in range((n + cc - 1) // cc):
                    acc += len(oracle.process_chunk("bpe", data[k * cc:(k + 1) * cc], om))
                    assert int(ends[k]) == acc, (variant, n, chunk, k)
            got = s.tokenize_host(data, chunk_size=65536)
            assert np.array_equal(got, oracle.run_buffer("bpe", data, 65536, 4, om))
            s.close()
    c.close()


@pytest.mark.parametrize("variant", [3, 4])
def test_fused_sweep_vs_oracle(nat, torch_mod, oracle, variant, monkeypatch):
    """The single-pass sweep (count + look-back + emit in one kernel, fused.cuh): sizes around its tile and round
    boundaries, ragged ends, every chunk size whose walls lie on tile boundaries, tables from no rule at all
    (T_out = N_in, the staging line's worst case) to every pair (one run per chunk), chunk walls on and inside
    tiles, and a chunk size shorter than a tile, which it must hand to the three-kernel sweep."""
    monkeypatch.setenv("BLT_SWEEP_VARIANT", str(variant))
    monkeypatch.setenv("BLT_DENSE", "0")
    tile = 30720 if variant == 3 else 23552   # 15 x 4 or 23 x 2 (worker warps x rounds) x 512 bytes
    c = nat.Context(0)
    rng = random.Random(3000 + variant)
    sizes = [1, 2, 15, 16, 17, 31, 33, 511, 512, 513, 4097, tile - 1, tile, tile + 1, tile + 15, tile + 16, tile + 17,
             2 * tile - 1, 2 * tile, 3 * tile + 511, 40 * tile + 7777, 301 * tile + 12345, 16 * MiB, 37 * MiB + 1]
    for n in sizes:
        for density in (1.0, 0.85, 0.3, 0.0):
            if n > 4 * MiB and density in (0.3,):
                continue
            alphabet = [97, 98, 99][: rng.choice([1, 2, 3])] if density == 1.0 else [97, 98, 99, 100, 32]
            data, pairs = _random_case(rng, n, alphabet, density)
            om = oracle.Merges(pairs)
            s = c.bpe_from_pairs(pairs)
            for chunk in (0, tile, 2 * tile, 8 * tile, 65536 + 16, 100001, 4 * MiB, 4000):
                if chunk > n and chunk != 0:
                    continue
                want = oracle.run_buffer("bpe", data, chunk or max(n, 1), 8, om)
                got, ends = resident(torch_mod, s, data, chunk)
                assert got.size == want.size and np.array_equal(got, want), (variant, n, density, chunk)
                cc = chunk or n
                nck = (n + cc - 1) // cc
                assert int(ends[-1]) == want.size
                if nck <= 48:
                    acc = 0
                    for k in range(nck):
                        acc += len(oracle.process_chunk("bpe", data[k * cc:(k + 1) * cc], om))
                        assert int(ends[k]) == acc, (variant, n, chunk, k)
            s.close()
    c.close()


@pytest.mark.parametrize("variant", [3, 4])
def test_fused_long_runs(nat, torch_mod, oracle, variant, monkeypatch):
    """Runs longer than a tile: every warp function and every tile function is the identity, so the carry has
    to travel through the look-back chain; a run start at an odd offset flips every parity behind it."""
    monkeypatch.setenv("BLT_SWEEP_VARIANT", str(variant))
    monkeypatch.setenv("BLT_DENSE", "0")
    c = nat.Context(0)
    pairs = {(97, 97): 256, (97, 98): 257, (98, 97): 258, (98, 98): 259}
    om = oracle.Merges(pairs)
    s = c.bpe_from_pairs(pairs)
    n = 6 * MiB + 5
    for head in (0, 1, 2, 3):
        data = np.full(n, 97, dtype=np.uint8)
        data[:head] = 99
        data[3 * MiB + 77] = 99
        for chunk in (0, 2 * MiB, 1 * MiB + 65536):
            got, _ = resident(torch_mod, s, data, chunk)
            assert np.array_equal(got, oracle.run_buffer("bpe", data, chunk or n, 4, om)), (head, chunk)
    ab = np.tile(np.frombuffer(b"ab", dtype=np.uint8), n // 2)
    got, _ = resident(torch_mod, s, ab, 1 * MiB)
    assert np.array_equal(got, oracle.run_buffer("bpe", ab, 1 * MiB, 4, om))
    # capacity: an output that is one token short must be reported, and nothing may be written behind it
    torch = torch_mod
    data = np.random.default_rng(5).choice(np.array([97, 98, 99], dtype=np.uint8), size=200001)
    want = oracle.run_buffer("bpe", data, 200001, 2, om)
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.full((want.size + 4096,), 0xA5, dtype=torch.uint8, device="cuda")
    with pytest.raises(nat.BltError) as ei:
        s.process_resident(d_in.data_ptr(), data.size, 0, d_out.data_ptr(), want.size - 2, 0, torch.cuda.current_stream().cuda_stream)
    assert ei.value.code == nat.ERR_CAPACITY
    assert bool((d_out[want.size - 2:] == 0xA5).all())
    ln = s.process_resident(d_in.data_ptr(), data.size, 0, d_out.data_ptr(), want.size, 0, torch.cuda.current_stream().cuda_stream)
    assert ln == want.size and np.array_equal(d_out[:ln].cpu().numpy(), want) and bool((d_out[want.size:] == 0xA5).all())
    s.close()
    c.close()


def test_dense_predictor_sequence(ctx, torch_mod, oracle):
    """One strategy, inputs that flip between merge-dense and sparse: whichever path the predictor picks
    (dense attempt, device-launched exact sweep, host-launched exact sweep), the bytes are the oracle's."""
    pairs = {(97, 98): 256, (98, 97): 257, (97, 97): 258, (98, 98): 259}
    om = oracle.Merges(pairs)
    s = ctx.bpe_from_pairs(pairs)
    rng = np.random.default_rng(7)
    n = 1 * MiB + 64
    dense_in = rng.integers(97, 99, size=n, dtype=np.uint8)             # every pair is a rule
    sparse_in = rng.integers(97, 101, size=n, dtype=np.uint8)           # c, d break the runs
    want = {True: oracle.run_buffer("bpe", dense_in, 65536, 4, om), False: oracle.run_buffer("bpe", sparse_in, 65536, 4, om)}
    plan = [True] * 3 + [False] * 40 + [True] * 40 + [False, True] * 10
    for i, d in enumerate(plan):
        got, _ = resident(torch_mod, s, dense_in if d else sparse_in, 65536)
        assert np.array_equal(got, want[d]), (i, d)
    s.close()


@pytest.mark.parametrize("variant", ["0", "1", "2", "3"])
def test_dense_prefix_and_exact_rest(nat, torch_mod, oracle, variant, monkeypatch):
    """The dense pass fails in some chunk: the chunks in front of it keep their dense output, the exact sweep redoes
    the rest (device-side launch).  Foreign bytes in the first / a middle / the last chunk, in several chunks, in the
    ragged tails, with chunk sizes that are and are not a whole number of 8 KiB work units; output and chunk ends are
    the oracle's."""
    monkeypatch.setenv("BLT_DENSE", "always")
    monkeypatch.setenv("BLT_SWEEP_VARIANT", variant)
    ctx = nat.Context(0)
    pairs = {(a, b): 256 + 4 * (a - 97) + (b - 97) for a in range(97, 101) for b in range(97, 101)}   # every pair of a..d
    om = oracle.Merges(pairs)
    s = ctx.bpe_from_pairs(pairs)
    rng = np.random.default_rng(23)
    for n, chunk in ((2 * MiB, 128 * 1024), (2 * MiB + 8192 + 37, 128 * 1024), (3 * MiB + 5, 192 * 1024 + 2), (1 * MiB, 0), (5 * MiB + 9, 1 * MiB)):
        c = chunk if chunk else n
        n_chunks = (n + c - 1) // c
        base = rng.integers(97, 101, size=n, dtype=np.uint8)
        spots = [[], [0], [n - 1], [n - 2], [c * (n_chunks // 2) + 1000], [c * (n_chunks // 2) + 1001],
                 [c - 2, c * (n_chunks - 1) + 4], [c * (n_chunks // 3) + 77, c * (n_chunks // 3) + 78, c * (2 * n_chunks // 3) + 12]]
        for sp in spots:
            data = base.copy()
            for q in sp:
                data[min(q, n - 1)] = 0x20
            want = oracle.run_buffer("bpe", data, c, 4, om)
            got, ends = resident(torch_mod, s, data, chunk)
            assert np.array_equal(got, want), (n, chunk, sp)
            want_ends = np.cumsum([len(oracle.process_chunk("bpe", data[k * c:(k + 1) * c], om)) for k in range(n_chunks)])
            assert np.array_equal(ends, want_ends), (n, chunk, sp)
    s.close()


def test_long_runs_and_carry_chains(ctx, torch_mod, oracle):
    """A whole chunk of one byte is a single run: parity must carry across threads, warps, tiles."""
    pairs = {(97, 97): 256, (97, 98): 257, (98, 97): 258, (98, 98): 259}
    om = oracle.Merges(pairs)
    s = ctx.bpe_from_pairs(pairs)
    n = 6 * MiB + 5
    for head in (0, 1, 2, 3):
        data = np.full(n, 97, dtype=np.uint8)
        data[:head] = 99                      # shifts the run start, flipping every parity after it
        data[3 * MiB + 77] = 99               # one break in the middle
        for chunk in (0, 2 * MiB, 2 * MiB + 1, 1 * MiB + 4097):
            got, _ = resident(torch_mod, s, data, chunk)
            assert np.array_equal(got, oracle.run_buffer("bpe", data, chunk or n, 4, om)), (head, chunk)
    ab = np.tile(np.frombuffer(b"ab", dtype=np.uint8), n // 2)
    got, _ = resident(torch_mod, s, ab, 1 * MiB)
    assert np.array_equal(got, oracle.run_buffer("bpe", ab, 1 * MiB, 4, om))


def test_general_maps_multi_sweep(ctx, torch_mod, oracle):
    """HashMap<(u16,u16),u16> beyond what merges.txt can express: chains, cycles, ids < 256."""
    rng = random.Random(5)
    cases = [
        ({(97, 98): 256, (256, 99): 257, (257, 100): 258}, b"abcd" * 5000 + b"abc"),
        ({(97, 97): 97}, b"a" * 100001),                               # cyclic, log2(n) sweeps
        ({(97, 98): 99, (99, 98): 100}, b"abb" * 3333),                # chain through ids < 256
        ({(120, 121): 90}, b"axyza" * 1000),
        ({(97, 98): 256, (256, 256): 257, (257, 257): 258, (258, 258): 259}, b"ab" * 40000),
    ]
    for _ in range(6):
        pool = [97, 98, 99, 256, 257, 258]
        md = {(rng.choice(pool), rng.choice(pool)): rng.choice(pool + [300, 301]) for _ in range(8)}
        if any(k[0] == v and k[1] == v for k, v in md.items()):
            md = {k: v for k, v in md.items() if not (k[0] == v and k[1] == v)}
        cases.append((md, bytes(rng.choice(b"abc") for _ in range(50000))))
    for md, raw in cases:
        data = np.frombuffer(raw, dtype=np.uint8)
        om = oracle.Merges(md)
        s = ctx.bpe_from_pairs(md)
        want, sweeps = oracle.process_chunk("bpe", data, om, want_sweeps=True)
        assert s.process_chunk(data).tobytes() == want, md
        for chunk in (0, 4096, 10000, 10001):
            got, ends = resident(torch_mod, s, data, chunk)
            assert got.tobytes() == bytes(oracle.run_buffer("bpe", data, chunk or data.size, 2, om)), (md, chunk)
            assert int(ends[(data.size + (chunk or data.size) - 1) // (chunk or data.size) - 1]) == got.size
        s.close()


def test_basic_random_and_ragged(ctx, torch_mod, oracle):
    from blt_b200 import synth
    s = ctx.basic()
    for n in (0, 1, 15, 16, 17, 4097, 1 * MiB + 7, 32 * MiB + 5):
        data = synth.random_bytes(n, 77 + n)
        want = oracle.run_buffer("basic", data, 1 << 22, 4)
        if n:
            got, ends = resident(torch_mod, s, data, 1 << 22)
            assert np.array_equal(got, want)
            assert int(ends[-1]) == 2 * n
        assert np.array_equal(s.tokenize_host(data, chunk_size=1 << 22), want)
        assert np.array_equal(s.process_chunk(data), np.frombuffer(oracle.process_chunk("basic", data), dtype=np.uint8))


def test_basic_output_aligned_to_16_not_32(ctx, torch_mod, oracle):
    """The C ABI promises 16-byte alignment only: an output pointer that is 16 (mod 32) must not take the
    32-byte store path (round-1 advisor finding: st.global.v8 faults on it, and the fault is sticky)."""
    from blt_b200 import synth
    torch = torch_mod
    s = ctx.basic()
    for n in (16, 4096 + 5, 3 * MiB + 17):
        data = synth.random_bytes(n, 123 + n)
        d_in = torch.from_numpy(data).cuda()
        d_buf = torch.full((2 * n + 64,), 0xA5, dtype=torch.uint8, device="cuda")
        off = 16 if d_buf.data_ptr() % 32 == 0 else 32 - (d_buf.data_ptr() % 32) + 16
        ln = s.process_resident(d_in.data_ptr(), n, 0, d_buf.data_ptr() + off, 2 * n, 0, torch.cuda.current_stream().cuda_stream)
        assert ln == 2 * n and (d_buf.data_ptr() + off) % 32 == 16
        h = d_buf.cpu().numpy()
        assert np.array_equal(h[off:off + 2 * n], oracle.run_buffer("basic", data, 1 << 22, 2))
        assert np.all(h[:off] == 0xA5) and np.all(h[off + 2 * n:] == 0xA5)


def test_strategy_outlives_its_context_handle(nat, oracle):
    """Context.close() before Strategy.close(): the strategy keeps the context alive (reference count) and still works."""
    c = nat.Context(0)
    pairs = {(97, 98): 256, (98, 99): 257}
    s = c.bpe_from_pairs(pairs)
    c.close()
    data = np.frombuffer(b"abcabcab" * 1000, dtype=np.uint8)
    assert np.array_equal(s.tokenize_host(data, chunk_size=4096), oracle.run_buffer("bpe", data, 4096, 2, oracle.Merges(pairs)))
    s.close()


def test_empty_and_capacity_errors(ctx, nat, torch_mod):
    s = ctx.bpe_from_pairs({(97, 98): 256})
    assert s.process_chunk(b"").size == 0                      # tokenizer.rs:57-59
    assert s.tokenize_host(b"", chunk_size=1024).size == 0     # pipeline.rs:103-105
    assert ctx.basic().tokenize_host(b"", 1024, nat.CONTENT_BIN).tobytes() == b"\xff\x03"
    small = np.empty(10, dtype=np.uint8)
    with pytest.raises(nat.BltError) as ei:
        s.process_chunk(b"x" * 1000, out=small)
    assert ei.value.code == nat.ERR_CAPACITY
    with pytest.raises(nat.BltError) as ei:                    # misaligned device pointer
        d = torch_mod.zeros(64, dtype=torch_mod.uint8, device="cuda")
        s.process_resident(d.data_ptr() + 1, 16, 0, d.data_ptr(), 32)
    assert ei.value.code == nat.ERR_INVALID_INPUT


def test_concurrent_process_chunk_on_one_strategy(ctx, oracle):
    """TokenizationStrategy is Send + Sync: up to num_threads calls at once (pipeline.rs:86-97)."""
    import threading
    from blt_b200 import synth
    data = synth.text(8 * MiB, 99)
    l, r = synth.merges_from_sample(data, 500)
    pairs = {(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))}
    s = ctx.bpe_from_pairs(pairs)
    om = oracle.Merges(pairs)
    want = [np.frombuffer(oracle.process_chunk("bpe", data[k * MiB:(k + 1) * MiB], om), dtype=np.uint8) for k in range(8)]
    errs = []

    def work(k):
        try:
            for _ in range(5):
                if not np.array_equal(s.process_chunk(data[k * MiB:(k + 1) * MiB]), want[k]):
                    errs.append(k)
        except Exception as e:  # pragma: no cover
            errs.append(repr(e))

    ts = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs


# ---- the five BASELINE.json configs ----------------------------------------------------------------------

def test_concurrent_fused_launches(nat, oracle, monkeypatch):
    """Eight host threads, each with its own pipe and stream, push 8 MiB chunks (274 tiles: a full grid of 148 CTAs)
    through the fused sweep at once.  Two fused launches sharing the SMs could wait for each other's unscheduled CTAs
    (the look-back needs every CTA of a launch resident); the launches of a device are chained on an event instead."""
    import threading
    from blt_b200 import synth
    monkeypatch.setenv("BLT_DENSE", "0")
    monkeypatch.setenv("BLT_SWEEP_VARIANT", "3")
    c = nat.Context(0)
    n_threads, chunk = 8, 8 * MiB
    data = synth.mixed(n_threads * chunk, 4242)
    l, r = synth.merges_from_sample(data, 4096)
    pairs = {(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))}
    s = c.bpe_from_pairs(pairs)
    om = oracle.Merges(pairs)
    want = [np.frombuffer(oracle.process_chunk("bpe", data[k * chunk:(k + 1) * chunk], om), dtype=np.uint8) for k in range(n_threads)]
    errs = []

    def work(k):
        try:
            for _ in range(6):
                if not np.array_equal(s.process_chunk(data[k * chunk:(k + 1) * chunk]), want[k]):
                    errs.append(("mismatch", k))
        except Exception as e:  # pragma: no cover
            errs.append((repr(e), k))

    ths = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs
    s.close()
    c.close()


def _check_chunks(torch, oracle, om, data, d_out_np, ends, chunk, which, mode="bpe"):
    for k in which:
        lo = 0 if k == 0 else int(ends[k - 1])
        want = np.frombuffer(oracle.process_chunk(mode, data[k * chunk:(k + 1) * chunk], om), dtype=np.uint8)
        assert np.array_equal(d_out_np[lo:int(ends[k])], want), k


def test_config1_basic_100mib(ctx, torch_mod, oracle):
    from blt_b200 import synth
    n = 100 * MiB
    data = synth.random_bytes(n, synth.SEED_CONFIG[1])
    got, ends = resident(torch_mod, ctx.basic(), data, 16 * MiB)
    assert got.size == 209715200
    assert np.array_equal(got, oracle.run_buffer("basic", data, 16 * MiB, 8))


def test_config2_bpe256_100mib(ctx, torch_mod, oracle, tmp_path):
    from blt_b200 import synth
    n = 100 * MiB
    data = synth.text(n, synth.SEED_CONFIG[2])
    l, r = synth.merges_from_sample(data, 256)
    synth.write_merges_file(str(tmp_path / "m.txt"), l, r)
    s = ctx.bpe_from_file(str(tmp_path / "m.txt"))
    om = oracle.Merges.from_file(str(tmp_path / "m.txt"))
    want = oracle.run_buffer("bpe", data, 16 * MiB, os.cpu_count() or 4, om)
    got, ends = resident(torch_mod, s, data, 16 * MiB)
    assert np.array_equal(got, want)
    assert np.array_equal(s.tokenize_host(data, chunk_size=16 * MiB), want)


@pytest.mark.parametrize("cfg", [3, 4])
def test_config3_and_4_one_gib(ctx, torch_mod, oracle, tmp_path, cfg):
    """1 GiB device-resident.  The WHOLE output and every chunk end are compared with the oracle's (sha256 of
    both, then the arrays), plus size-independent properties (decode back to the chunk's bytes, idempotence)."""
    from blt_b200 import synth
    torch = torch_mod
    n, chunk = 1 << 30, 16 * MiB
    if cfg == 3:
        data = synth.text(n, synth.SEED_CONFIG[3])
        l, r = synth.merges_from_sample(data, 32768)
        synth.write_merges_file(str(tmp_path / "m.txt"), l, r)
        s = ctx.bpe_from_file(str(tmp_path / "m.txt"))
        om = oracle.Merges.from_file(str(tmp_path / "m.txt"))
        assert s.num_merges == 32768 and max(om.to_dict().values()) == 33023
    else:
        data = synth.adversarial(n, synth.SEED_CONFIG[4])
        pairs = {p: 256 + i for i, p in enumerate(synth.adversarial_pairs())}
        s = ctx.bpe_from_pairs(pairs)
        om = oracle.Merges(pairs)
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    d_ends = torch.zeros(n // chunk, dtype=torch.int64, device="cuda")
    out_len = s.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), 2 * n, d_ends.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
    ends = d_ends.cpu().numpy()
    out = d_out[:out_len].cpu().numpy()
    assert int(ends[-1]) == out_len and np.all(np.diff(ends) > 0)
    want = oracle.run_buffer("bpe", data, chunk, os.cpu_count() or 4, om)
    import hashlib
    assert hashlib.sha256(out.tobytes()).hexdigest() == hashlib.sha256(want.tobytes()).hexdigest()
    assert out.size == want.size and np.array_equal(out, want)
    _check_chunks(torch, oracle, om, data, out, ends, chunk, [0, 31, 63])       # and the chunk ends sit where the oracle's chunks end
    # property: every chunk's tokens decode back to exactly that chunk's bytes
    first_of = np.arange(65536, dtype=np.int64)
    second_of = np.zeros(65536, dtype=np.int64)
    for (a, b), v in om.to_dict().items():
        first_of[v], second_of[v] = a, b
    for k in (5, 40):
        lo = 0 if k == 0 else int(ends[k - 1])
        toks = out[lo:int(ends[k])].view(">u2").astype(np.int64)
        merged = toks >= 256
        width = 1 + merged.astype(np.int64)
        assert int(width.sum()) == chunk                          # each merge removes exactly one symbol
        pos = np.cumsum(width) - width
        rec = np.zeros(chunk, dtype=np.uint8)
        rec[pos] = first_of[toks]
        rec[pos[merged] + 1] = second_of[toks[merged]]
        assert np.array_equal(rec, data[k * chunk:(k + 1) * chunk]), k
    # property: idempotence of the launch (same input, same table -> same bytes)
    d_out2 = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    out_len2 = s.process_resident(d_in.data_ptr(), n, chunk, d_out2.data_ptr(), 2 * n, 0, torch.cuda.current_stream().cuda_stream)
    assert out_len2 == out_len and torch.equal(d_out2[:out_len2], d_out[:out_len])
    if cfg == 4:   # chain maps through the pair API on a 64 MiB slice
        sl = data[: 64 * MiB]
        md = {(97, 98): 99, (99, 98): 100, (97, 97): 256, (256, 256): 257, (100, 100): 97}
        sg = ctx.bpe_from_pairs(md)
        got, _ = resident(torch, sg, sl, chunk)
        assert np.array_equal(got, oracle.run_buffer("bpe", sl, chunk, os.cpu_count() or 4, oracle.Merges(md)))


@pytest.mark.parametrize("variant", [0, 3])
def test_config4_sparse_table_runs_end_inside_chunks(nat, torch_mod, oracle, variant, monkeypatch):
    """Config 4 with a table that omits pairs ((a,a) and (a,b) only): the dense pass can never hold, runs of every
    listed length (31/32/33, 2^20 +- 1, 16 MiB + 1, ...) start at arbitrary offsets and end INSIDE 16 MiB chunks or
    straddle their walls, so run parity, ties and the carry across tiles are exercised at BASELINE size (1 GiB)."""
    monkeypatch.setenv("BLT_SWEEP_VARIANT", str(variant))
    torch = torch_mod
    n, chunk = 1 << 30, 16 * MiB
    rng = np.random.default_rng(4040 + variant)
    data = np.empty(n, dtype=np.uint8)
    lens = [1, 2, 3, 15, 16, 17, 31, 32, 33, 255, 256, 257, 511, 512, 513, 30719, 30720, 30721, (1 << 20) - 1, 1 << 20, (1 << 20) + 1,
            16 * MiB - 1, 16 * MiB, 16 * MiB + 1]
    pos = 0
    seps = np.frombuffer(b"bcdb", dtype=np.uint8)
    while pos < n:
        ln = int(lens[rng.integers(len(lens))]) if rng.random() < 0.9 else int(rng.integers(1, 70000))
        ln = min(ln, n - pos)
        data[pos:pos + ln] = 97
        pos += ln
        k = min(int(rng.integers(1, 4)), n - pos)      # 1-3 separator bytes: b (so that (a,b) fires), c, d
        if k > 0:
            data[pos:pos + k] = seps[rng.integers(0, 3, size=k)]
            pos += k
    pairs = {(97, 97): 256, (97, 98): 257}
    c = nat.Context(0)
    s = c.bpe_from_pairs(pairs)
    om = oracle.Merges(pairs)
    want = oracle.run_buffer("bpe", data, chunk, os.cpu_count() or 4, om)
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    d_ends = torch.zeros(n // chunk, dtype=torch.int64, device="cuda")
    for rep in range(2):    # the second call goes through the predictor's "exact directly" path
        out_len = s.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), 2 * n, d_ends.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream)
        assert out_len == want.size
        assert torch.equal(d_out[:out_len], torch.from_numpy(want).cuda()), rep
    ends = d_ends.cpu().numpy()
    _check_chunks(torch, oracle, om, data, d_out[:out_len].cpu().numpy(), ends, chunk, [0, 1, 2, 33, 63])
    s.close()
    c.close()


def test_config5_file_to_file(nat, oracle, tmp_path_factory):
    """BASELINE configs[4]: 8 GiB synthetic corpus, 60 000 merges, chunks sharded over 1 GPU and over all visible
    GPUs, file to file (pipeline.rs:56-131); the output file must equal the oracle's (ora_run_files) byte for byte.
    Scaled down (and says so) only when the box lacks the ~45 GiB of RAM / tmpfs the full size needs."""
    import hashlib
    from blt_b200 import synth
    d = "/dev/shm" if os.path.isdir("/dev/shm") else str(tmp_path_factory.mktemp("cfg5"))
    st = os.statvfs(d)
    free = st.f_bavail * st.f_frsize
    try:
        import psutil
        free = min(free, psutil.virtual_memory().available)
    except ImportError:
        pass
    n = 8 << 30
    while n > (1 << 30) and free < 5 * n + (4 << 30):
        n >>= 1
    if n != 8 << 30:
        print(f"test_config5_file_to_file: scaled down to {n >> 30} GiB ({free >> 30} GiB free)")
    inp, outp, ref, mp = (os.path.join(d, f"blt_t5_{os.getpid()}.{e}") for e in ("in", "out", "ref", "merges.txt"))
    try:
        data = synth.text(n, synth.SEED_CONFIG[5])
        l, r = synth.merges_from_sample(data, 60000)
        synth.write_merges_file(mp, l, r)
        data.tofile(inp)
        del data
        oracle.run_files("bpe", inp, ref, 16 * MiB, os.cpu_count() or 4, oracle.Merges.from_file(mp))

        def sha(path):
            h = hashlib.sha256()
            with open(path, "rb") as f:
                for blk in iter(lambda: f.read(64 << 20), b""):
                    h.update(blk)
            return h.hexdigest()

        want = sha(ref)
        want_size = os.path.getsize(ref)
        os.unlink(ref)
        for g in sorted({1, nat.device_count()}):
            nat.run_tokenizer(inp, outp, merges_file=mp, chunk_size="16MB", num_gpus=g)
            assert os.path.getsize(outp) == want_size, g
            assert sha(outp) == want, g
    finally:
        for f in (inp, outp, ref, mp):
            if os.path.exists(f):
                os.unlink(f)


# ---- file to file: CLI and Python binding -------------------------------------------------------------------

def test_resident_beyond_4gib(ctx, torch_mod, oracle):
    """More than 2^32 input bytes in one launch (every index on the device path is 64-bit): basic, the
    exact sweep (sparse table) and the dense pass (full table), checked on the chunks around the 4 GiB
    mark and on the ragged last chunk."""
    from blt_b200 import synth
    torch = torch_mod
    chunk = 16 * MiB
    n = (4 << 30) + 3 * chunk + 12345
    n_chunks = (n + chunk - 1) // chunk
    data = synth.text(n, 0xB170099)
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.empty(2 * n + 16, dtype=torch.uint8, device="cuda")
    d_ends = torch.zeros(n_chunks, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    which = [0, 127, 254, 255, 256, 257, n_chunks - 2, n_chunks - 1]

    def run(strat):
        out_len = strat.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), 2 * n, d_ends.data_ptr(), stream)
        ends = d_ends.cpu().numpy()
        assert int(ends[-1]) == out_len
        return out_len, ends

    def chunk_bytes(ends, k):
        lo = 0 if k == 0 else int(ends[k - 1])
        return d_out[lo:int(ends[k])].cpu().numpy()

    out_len, ends = run(ctx.basic())
    assert out_len == 2 * n
    for k in which:
        want = np.frombuffer(oracle.process_chunk("basic", data[k * chunk:(k + 1) * chunk], None), dtype=np.uint8)
        assert np.array_equal(chunk_bytes(ends, k), want), ("basic", k)
    for rules in (256, 32768):
        l, r = synth.merges_from_sample(data[: 64 * MiB], rules)
        pairs = {(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))}
        om = oracle.Merges(pairs)
        s = ctx.bpe_from_pairs(pairs)
        out_len, ends = run(s)
        assert np.all(np.diff(ends) > 0)
        for k in which:
            want = np.frombuffer(oracle.process_chunk("bpe", data[k * chunk:(k + 1) * chunk], om), dtype=np.uint8)
            assert np.array_equal(chunk_bytes(ends, k), want), (rules, k)
        s.close()


def _resident_detok(torch, strat, toks: np.ndarray):
    d_in = torch.from_numpy(np.ascontiguousarray(toks)).cuda() if toks.size else torch.empty(16, dtype=torch.uint8, device="cuda")
    d_out = torch.full((toks.size + 32,), 0xEE, dtype=torch.uint8, device="cuda")
    n = strat.detokenize_resident(d_in.data_ptr(), toks.size, d_out.data_ptr(), toks.size, torch.cuda.current_stream().cuda_stream)
    assert bool((d_out[toks.size:] == 0xEE).all()), "wrote past the capacity"
    return d_out[:n].cpu().numpy()


@pytest.mark.parametrize("variant", ["0", "1"])
def test_detokenize_vs_oracle(nat, torch_mod, oracle, variant, monkeypatch):
    """GPU detokenizer against the oracle's on arbitrary valid token streams, ragged sizes,
    contiguous ids and ids with holes, device-resident and host entry points; the count/scan/emit form (0, the
    default) and the fused single pass (1)."""
    monkeypatch.setenv("BLT_DETOK_VARIANT", variant)
    ctx = nat.Context(0)
    rng = np.random.default_rng(11)
    tables = {
        "contiguous": {(int(k) & 255, int(k) >> 8): 256 + i for i, k in enumerate(rng.choice(65536, 3000, replace=False))},
        "holes": {(97 + i % 5, 97 + i // 5): 300 + 37 * i for i in range(25)},
        "top": {(1, 2): 65535, (3, 4): 256},
    }
    for name, pairs in tables.items():
        # later duplicates of a key overwrite: keep the final map, and drop ids that collide
        final = {}
        for k, v in pairs.items():
            final[k] = v
        ids = np.array(sorted(set(final.values())), dtype=np.int64)
        if len(ids) != len(final):
            continue
        om = oracle.Merges(final)
        s = ctx.bpe_from_pairs(final)
        for n_tok in (0, 1, 2, 7, 8, 9, 255, 256, 257, 4097, 32767, 32768, 32769, 65536, 100001, 1 * MiB + 3, 3 * MiB + 5,
                      148 * 32768 * 2 + 11):
            for p_wide in (0.0, 0.5, 1.0, 0.03, 0.97):
                wide = rng.random(n_tok) < p_wide
                toks = np.where(wide, ids[rng.integers(0, len(ids), n_tok)], rng.integers(0, 256, n_tok)).astype(">u2")
                stream = toks.view(np.uint8)
                want = oracle.detokenize(stream, om)
                assert np.array_equal(_resident_detok(torch_mod, s, stream), want), (name, n_tok, p_wide)
            assert np.array_equal(s.detokenize_host(stream), want), (name, n_tok)
        s.close()
    b = ctx.basic()
    raw = rng.integers(0, 256, 1 * MiB + 77).astype(np.uint8)
    assert np.array_equal(b.detokenize_host(b.tokenize_host(raw, chunk_size=65536)), raw)
    p = ctx.passthrough()
    assert np.array_equal(p.detokenize_host(raw[: 2 * (raw.size // 2)]), raw[: 2 * (raw.size // 2)])


@pytest.mark.parametrize("variant", ["0", "1"])
def test_detokenize_errors_and_prefix(nat, torch_mod, oracle, variant, monkeypatch):
    monkeypatch.setenv("BLT_DETOK_VARIANT", variant)
    ctx = nat.Context(0)
    pairs = {(97, 98): 256, (98, 97): 300}
    s = ctx.bpe_from_pairs(pairs)
    # an output that is too small, far into a long stream: a capacity error, nothing written behind the capacity
    torch = torch_mod
    long_toks = np.tile(np.frombuffer(b"\x01\x00\x00c", dtype=np.uint8), 3 * MiB)      # (a,b), c -> 3 bytes per 2 tokens
    d_in = torch.from_numpy(long_toks).cuda()
    cap = 5 * MiB
    d_out = torch.full((cap + 4096,), 0xEE, dtype=torch.uint8, device="cuda")
    with pytest.raises(nat.BltError) as e:
        s.detokenize_resident(d_in.data_ptr(), long_toks.size, d_out.data_ptr(), cap, torch.cuda.current_stream().cuda_stream)
    assert e.value.code == -7
    assert bool((d_out[cap:] == 0xEE).all()), "wrote past the capacity"
    ok = np.frombuffer(b"\xff\x01\x01\x00\x00c\x01\x2c", dtype=np.uint8)          # Text prefix, (a,b), c, (b,a)
    assert bytes(s.detokenize_host(ok, has_content_type=True)) == b"abcba"
    for bad in (b"\x00", b"\x01\x01", b"\x01\x00" * 5000 + b"\x01\x2b", b"\xff\x01\x00a"):  # odd, unknown ids (hole, 299), prefix unasked
        with pytest.raises(nat.BltError) as e:
            s.detokenize_host(np.frombuffer(bad, dtype=np.uint8))
        assert e.value.code == -3, bad[:8]
    with pytest.raises(nat.BltError) as e:
        s.detokenize_host(np.frombuffer(b"\x00a", dtype=np.uint8), has_content_type=True)
    assert e.value.code == -3
    with pytest.raises(nat.BltError) as e:                                              # capacity
        s.detokenize_host(np.frombuffer(b"\x01\x00" * 64, dtype=np.uint8), out=np.empty(100, dtype=np.uint8))
    assert e.value.code == -7
    s.close()
    for not_invertible in ({(97, 98): 256, (98, 97): 256}, {(97, 98): 256, (256, 99): 257}, {(97, 97): 97}):
        g = ctx.bpe_from_pairs(not_invertible)
        with pytest.raises(nat.BltError) as e:
            g.detokenize_host(np.frombuffer(b"\x00a", dtype=np.uint8))
        assert e.value.code == -2
        g.close()


@pytest.mark.parametrize("cfg", [2, 3])
def test_round_trip_one_gib(ctx, torch_mod, cfg):
    """Size-independent property at BASELINE size: detokenize(tokenize(x)) == x for the whole GiB, on the
    device (sparse table -> exact sweep, full table -> dense pass)."""
    from blt_b200 import synth
    torch = torch_mod
    n, chunk = 1 << 30, 16 * MiB
    data = synth.text(n, synth.SEED_CONFIG[cfg])
    l, r = synth.merges_from_sample(data, 256 if cfg == 2 else 32768)
    s = ctx.bpe_from_pairs({(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))})
    d_in = torch.from_numpy(data).cuda()
    d_tok = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    n_tok_bytes = s.process_resident(d_in.data_ptr(), n, chunk, d_tok.data_ptr(), 2 * n, 0, stream)
    d_back = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    n_back = s.detokenize_resident(d_tok.data_ptr(), n_tok_bytes, d_back.data_ptr(), n, stream)
    assert n_back == n
    assert torch.equal(d_back[:n], d_in)
    s.close()


def _pair_counts_numpy(data: np.ndarray) -> np.ndarray:
    if data.size < 2:
        return np.zeros(65536, dtype=np.uint64)
    keys = (data[:-1].astype(np.uint32) << 8) | data[1:]
    return np.bincount(keys, minlength=65536).astype(np.uint64)


def test_pair_histogram_and_training(ctx, torch_mod):
    """the GPU pair histogram against numpy, and the tables it yields against the workload
    generator's CPU rule (most frequent first, ties by b0*256+b1, padded with unobserved pairs)."""
    from blt_b200 import synth
    rng = np.random.default_rng(5)
    cases = [np.zeros(0, np.uint8), np.array([7], np.uint8), np.array([7, 9], np.uint8), rng.integers(0, 256, 17, dtype=np.uint8),
             rng.integers(0, 256, 16 * 1024 * 3 + 1, dtype=np.uint8), rng.integers(0, 256, 1 * MiB + 3, dtype=np.uint8),
             np.zeros(300000, np.uint8),                      # one pair 299 999 times: u16 counters must be flushed in time
             np.tile(np.frombuffer(b"ab", np.uint8), 100001), synth.text(40 * MiB + 5, 77)]
    for data in cases:
        assert np.array_equal(ctx.count_pairs(data), _pair_counts_numpy(data)), data.size
    # device-resident entry point
    data = cases[-1]
    d_in = torch_mod.from_numpy(data).cuda()
    d_counts = torch_mod.full((65536,), 123, dtype=torch_mod.int64, device="cuda")
    ctx.count_pairs_resident(d_in.data_ptr(), data.size, d_counts.data_ptr(), torch_mod.cuda.current_stream().cuda_stream)
    assert np.array_equal(d_counts.cpu().numpy().astype(np.uint64), _pair_counts_numpy(data))
    # training == the workload generator's rule on the same sample
    sample = data[: 16 * MiB]
    for k, pad in ((256, False), (32768, True)):
        l, r = ctx.train_merges(sample, k, pad)
        wl, wr = synth.merges_from_sample(sample, k)
        assert np.array_equal(l, wl) and np.array_equal(r, wr), k


@pytest.mark.parametrize("variant,dense", [(0, "0"), (1, "0"), (2, "0"), (0, "always"), (3, "0"), (4, "0"), (3, "always")])
def test_no_writes_outside_the_buffers(nat, torch_mod, oracle, variant, dense, monkeypatch):
    """Input, output (capacity exactly 2n) and chunk_ends (exactly one entry per chunk) sit between canaries
    in one allocation: ragged sizes and tiny chunks must leave every canary and the input intact.  (Found by
    tools/fuzz_gpu.py: chunk walls behind the end of the input used to write one entry past chunk_ends.)"""
    torch = torch_mod
    monkeypatch.setenv("BLT_SWEEP_VARIANT", str(variant))
    monkeypatch.setenv("BLT_DENSE", dense)
    c = nat.Context(0)
    rng = np.random.default_rng(3)
    pairs = {(97, 97): 256, (97, 98): 257, (98, 97): 258, (98, 98): 259, (99, 97): 260, (97, 99): 261, (99, 99): 262}
    om = oracle.Merges(pairs)
    s = c.bpe_from_pairs(pairs)
    G = 4096
    al = lambda x: (x + 255) // 256 * 256
    stream = torch.cuda.current_stream().cuda_stream
    for n, chunk in ((6, 0), (15, 0), (21, 0), (35, 16), (253, 2), (253, 16), (2119, 16), (4095, 2047), (4097, 4096),
                     (2 * 4096 - 5, 4096), (70001, 1000), (1 * MiB + 9, 65536),
                     (32768, 0), (32769, 32768), (3 * 32768 + 17, 32768), (5 * 23552 - 1, 23552), (7 * 30720 + 3, 30720), (1 * MiB + 9, 0)):
        data = rng.choice(np.array([97, 98, 99], dtype=np.uint8), size=n)
        eff = chunk if chunk and chunk < n else n
        nc = (n + eff - 1) // eff
        want = oracle.run_buffer("bpe", data, eff, 2, om)
        o_in = G; o_out = o_in + al(n) + G; o_ends = o_out + al(2 * n) + G; total = o_ends + al(8 * nc) + G
        buf = torch.full((total,), 0xA5, dtype=torch.uint8, device="cuda")
        buf[o_in:o_in + n] = torch.from_numpy(data).cuda()
        base = buf.data_ptr()
        for rep in range(2):
            ln = s.process_resident(base + o_in, n, chunk, base + o_out, 2 * n, base + o_ends, stream)
            h = buf.cpu().numpy()
            assert np.array_equal(h[o_in:o_in + n], data), (n, chunk, rep, "input modified")
            for lo, hi in ((0, o_in), (o_in + n, o_out), (o_out + 2 * n, o_ends), (o_ends + 8 * nc, total)):
                assert np.all(h[lo:hi] == 0xA5), (n, chunk, rep, "canary", lo)
            assert np.array_equal(h[o_out:o_out + ln], want), (n, chunk, rep)
            assert int(h[o_ends:o_ends + 8 * nc].view(np.int64)[-1]) == want.size
    s.close()
    c.close()


def test_fuzz_short():
    """A bounded run of the randomised parity fuzzer (tools/fuzz_gpu.py) with a fixed seed."""
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_gpu.py"), "--seconds", "25", "--seed", "424242"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_cli_file_to_file_and_stdin(oracle, tmp_path):
    from blt_b200 import synth
    data = synth.text(5 * MiB + 321, 4242)
    (tmp_path / "in.bin").write_bytes(data.tobytes())
    l, r = synth.merges_from_sample(data, 300)
    synth.write_merges_file(str(tmp_path / "m.txt"), l, r)
    om = oracle.Merges.from_file(str(tmp_path / "m.txt"))
    # tests/cli.rs:46-80 (file -> file, basic), with default chunking
    assert subprocess.run([BLT, "--input", str(tmp_path / "in.bin"), "--output", str(tmp_path / "o1.bin")]).returncode == 0
    assert (tmp_path / "o1.bin").read_bytes() == bytes(oracle.run_buffer("basic", data, 16 * MiB, 4))
    # BPE, explicit chunk size (clamped up to 256 KiB, chunking.rs:29), content-type prefix
    for cs, eff in (("1KB", 256 * 1024), ("1MB", MiB), ("300000", 300000)):
        assert subprocess.run([BLT, "-i", str(tmp_path / "in.bin"), "-o", str(tmp_path / "o2.bin"), "--merges",
                               str(tmp_path / "m.txt"), "--chunksize", cs, "--type", "text", "--threads", "2"]).returncode == 0
        assert (tmp_path / "o2.bin").read_bytes() == bytes(oracle.run_buffer("bpe", data, eff, 4, om, 0xFF01)), cs
    # stdin -> stdout rows of tests/cli.rs
    r = subprocess.run([BLT], input=b"hello world", capture_output=True)
    assert r.returncode == 0 and r.stdout == bytes(oracle.run_buffer("basic", b"hello world", 1 << 20, 1))
    r = subprocess.run([BLT, "--type", "text"], input=b"test", capture_output=True)
    assert r.stdout == b"\xff\x01\x00t\x00e\x00s\x00t"
    (tmp_path / "ab.txt").write_text("97 98\n")
    r = subprocess.run([BLT, "--merges", str(tmp_path / "ab.txt")], input=b"ab c ab", capture_output=True)
    assert r.stdout == be([256, 32, 99, 32, 256])
    r = subprocess.run([BLT, "--chunksize", "1KB"], input=b"some data", capture_output=True)
    assert r.stdout == bytes(oracle.run_buffer("basic", b"some data", 1 << 20, 1))
    # --detokenize (an addition): the CLI undoes its own output, file to file and over pipes
    assert subprocess.run([BLT, "--detokenize", "-i", str(tmp_path / "o2.bin"), "-o", str(tmp_path / "back.bin"), "--merges",
                           str(tmp_path / "m.txt"), "--type", "text"]).returncode == 0
    assert (tmp_path / "back.bin").read_bytes() == data.tobytes()
    r = subprocess.run([BLT, "--detokenize", "--merges", str(tmp_path / "ab.txt")], input=be([256, 32, 99, 32, 256]), capture_output=True)
    assert r.returncode == 0 and r.stdout == b"ab c ab"
    r = subprocess.run([BLT, "--detokenize", "--merges", str(tmp_path / "ab.txt")], input=be([300]), capture_output=True)
    assert r.returncode == 1 and b"not in the table" in r.stderr
    # empty file -> empty output
    (tmp_path / "empty").write_bytes(b"")
    assert subprocess.run([BLT, "-i", str(tmp_path / "empty"), "-o", str(tmp_path / "o3.bin"), "--merges", str(tmp_path / "m.txt")]).returncode == 0
    assert (tmp_path / "o3.bin").read_bytes() == b""


def test_output_that_cannot_grow_is_an_io_error(tmp_path):
    """a full filesystem must surface as an error (the reference returns io::Error and exits 1), never as
    a SIGBUS from the output mapping.  A file-size limit makes the output's ftruncate / fallocate / pwrite fail with
    EFBIG: the CLI must exit 1 with a message, and the same run without the limit must succeed."""
    import resource
    import signal
    blt = os.path.join(ROOT, "blt_b200", "lib", "blt")
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    rng = np.random.default_rng(7)
    rng.integers(0, 256, size=4 * MiB, dtype=np.uint8).tofile(inp)

    def limited():
        signal.signal(signal.SIGXFSZ, signal.SIG_IGN)
        resource.setrlimit(resource.RLIMIT_FSIZE, (1 * MiB, 1 * MiB))

    r = subprocess.run([blt, "-i", inp, "-o", outp, "--chunksize", "1MB"], capture_output=True, text=True, preexec_fn=limited, timeout=300)
    assert r.returncode == 1, (r.returncode, r.stderr[-500:])      # not -SIGBUS, not 0
    assert r.stderr.strip() != ""
    r = subprocess.run([blt, "-i", inp, "-o", outp, "--chunksize", "1MB"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    assert os.path.getsize(outp) == 8 * MiB


def test_python_bytetokenizer(oracle, tmp_path):
    """The value-free smoke tests of blt_python/tests/test_tokenizer.py, with the values pinned."""
    import blt_b200
    (tmp_path / "in.bin").write_bytes(b"hello world")
    blt_b200.ByteTokenizer().tokenize_file(str(tmp_path / "in.bin"), str(tmp_path / "out.bin"))
    assert (tmp_path / "out.bin").read_bytes() == bytes(oracle.run_buffer("basic", b"hello world", 1 << 20, 1))
    (tmp_path / "in.bin").write_bytes(b"ab")
    blt_b200.ByteTokenizer(merges={(97, 98): 256}).tokenize_file(str(tmp_path / "in.bin"), str(tmp_path / "out.bin"))
    assert (tmp_path / "out.bin").read_bytes() == b"\x01\x00"
    (tmp_path / "in.bin").write_bytes(b"")
    blt_b200.ByteTokenizer().tokenize_file(str(tmp_path / "in.bin"), str(tmp_path / "out.bin"))
    assert (tmp_path / "out.bin").read_bytes() == b""
    big = b"x" * (100 * 1024)
    (tmp_path / "in.bin").write_bytes(big)
    t = blt_b200.ByteTokenizer(threads=2, chunk_size="1MB", memory_cap=50, content_type="Bin")
    t.tokenize_file(str(tmp_path / "in.bin"), str(tmp_path / "out.bin"))
    assert (tmp_path / "out.bin").read_bytes() == b"\xff\x03" + bytes(oracle.run_buffer("basic", big, 1 << 20, 1))
    t.detokenize_file(str(tmp_path / "out.bin"), str(tmp_path / "back.bin"))            # an addition: the inverse
    assert (tmp_path / "back.bin").read_bytes() == big
    with pytest.raises(FileNotFoundError):
        blt_b200.ByteTokenizer().tokenize_file(str(tmp_path / "nope"), str(tmp_path / "out.bin"))
