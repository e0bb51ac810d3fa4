"""Pins the CPU oracle (oracle/blt_oracle.cpp) and the independent Python model on every golden
vector the reference's own tests hold for this path (tests/golden/reference_vectors.json, each row
citing reference file:line), then cross-checks the two restatements against each other and against
the closed parallel form the CUDA kernels implement.  No GPU needed."""
import json
import os
import random

import pytest
from hypothesis import given, settings, strategies as st

from oracle import py_model as pm

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))
DER = json.load(open(os.path.join(HERE, "golden", "derived_vectors.json")))


def be(tokens):
    return b"".join(int(t).to_bytes(2, "big") for t in tokens)


def merges_dict(rows):
    return {(a, b): v for a, b, v in rows}


# ---- reference-held vectors ---------------------------------------------------------------------

@pytest.mark.parametrize("row", REF["bpe"], ids=lambda r: r["src"])
def test_ref_bpe_vectors(oracle, row, tmp_path):
    data = row["input"].encode()
    if "merges_file" in row:
        p = tmp_path / "merges.txt"
        p.write_text(row["merges_file"])
        m = oracle.Merges.from_file(str(p))
        md = m.to_dict()
    else:
        md = merges_dict(row["merges"])
        m = oracle.Merges(md)
    assert oracle.process_chunk("bpe", data, m) == be(row["tokens"])
    assert pm.bpe_chunk(data, md) == be(row["tokens"])
    # same answer through the whole in-memory pipeline with one chunk
    assert bytes(oracle.run_buffer("bpe", data, 1 << 20, 2, m)) == be(row["tokens"])


@pytest.mark.parametrize("row", REF["basic"], ids=lambda r: r["src"])
def test_ref_basic_vectors(oracle, row):
    data = row["input"].encode()
    want = bytes.fromhex(row["bytes_hex"])
    assert oracle.process_chunk("basic", data) == want
    assert pm.basic_chunk(data) == want
    chunk = oracle.effective_chunk_size(oracle.parse_chunk_size(row["chunksize"]), 4, 80, 1 << 36) \
        if "chunksize" in row else 1 << 20
    assert bytes(oracle.run_buffer("basic", data, chunk, row.get("threads", 4))) == want


@pytest.mark.parametrize("row", REF["passthrough"], ids=lambda r: r["src"])
def test_ref_passthrough_vectors(oracle, row):
    data = row["input"].encode()
    want = bytes.fromhex(row["bytes_hex"])
    assert oracle.process_chunk("passthrough", data) == want
    assert bytes(oracle.run_buffer("passthrough", data, 1 << 20, 1)) == want


def test_ref_content_type(oracle):
    t = REF["content_type_tokens"]
    assert [oracle.content_type_token(i) for i in range(4)] == [t["text"], t["audio"], t["bin"], t["video"]]
    row = REF["content_type"][0]
    got = oracle.run_buffer("basic", row["input"].encode(), 1 << 20, 1, content_type_token=t[row["type"]])
    assert bytes(got) == bytes.fromhex(row["bytes_hex"])
    assert pm.run_buffer("basic", row["input"].encode(), 1 << 20, None, t[row["type"]]) == bytes.fromhex(row["bytes_hex"])


@pytest.mark.parametrize("row", REF["merges_files"], ids=lambda r: r["src"])
def test_ref_merges_files(oracle, row, tmp_path):
    p = tmp_path / "m.txt"
    p.write_text(row["text"])
    assert oracle.Merges.from_file(str(p)).to_dict() == merges_dict(row["map"])
    assert pm.load_bpe_merges_text(row["text"]) == merges_dict(row["map"])


@pytest.mark.parametrize("row", REF["merges_file_errors"], ids=lambda r: r["src"])
def test_ref_merges_file_errors(oracle, row, tmp_path):
    kinds = {"InvalidData": oracle.ORA_INVALID_DATA, "NotFound": oracle.ORA_NOT_FOUND}
    p = tmp_path / "m.txt"
    if not row.get("missing_file"):
        p.write_text(row["text"])
    with pytest.raises(oracle.OracleError) as ei:
        oracle.Merges.from_file(str(p))
    assert ei.value.kind == kinds[row["kind"]]
    if "contains" in row:
        assert row["contains"] in ei.value.message


def test_ref_chunk_size_parse(oracle):
    for s, v in REF["chunk_size_parse"]["valid"]:
        assert oracle.parse_chunk_size(s) == v
    for s in REF["chunk_size_parse"]["invalid"]:
        with pytest.raises(oracle.OracleError) as ei:
            oracle.parse_chunk_size(s)
        assert ei.value.kind == oracle.ORA_INVALID_INPUT


def test_ref_chunk_size_clamp_and_threads(oracle):
    for cli, want in REF["chunk_size_clamp"]["cases"]:
        assert oracle.effective_chunk_size(cli, 4, 80, 64 << 30) == want
    for v, want in REF["thread_count"]["cases"]:
        assert oracle.determine_thread_count(v, 8) == want
    assert oracle.determine_thread_count(None, 8) == 8
    assert oracle.determine_thread_count(None, 0) == 1
    # dynamic path, chunking.rs:33-61: bounds the reference test asserts, plus exact arithmetic
    assert oracle.effective_chunk_size(None, 4, 80, 64 << 30) == 16 << 20
    assert oracle.effective_chunk_size(None, 4, 1, 1 << 30) == 1 << 20            # clamps up to 1 MiB
    assert oracle.effective_chunk_size(None, 128, 80, 8 << 30) == int((8 << 30) * 0.8) // 128 // 4
    assert oracle.effective_chunk_size(None, 128, 80, 4 << 30) == int((4 << 30) * 0.8) // 128 // 4


# ---- merges loader corner cases derived from config_loader.rs ---------------------------------------

def test_loader_corner_cases(oracle, tmp_path):
    def load(text, binary=False):
        p = tmp_path / "c.txt"
        p.write_bytes(text if binary else text.encode())
        return oracle.Merges.from_file(str(p)).to_dict()

    assert load("97 98") == {(97, 98): 256}                       # no trailing newline
    assert load("97 98\r\n99 100\r\n") == {(97, 98): 256, (99, 100): 257}   # CRLF stripped by lines()
    assert load("  97\t 98  \n") == {(97, 98): 256}               # split_whitespace
    assert load("+97 098\n") == {(97, 98): 256}                   # u8::from_str accepts + and zeros
    assert load("\n\n#x\n0 0\n255 255\n") == {(0, 0): 256, (255, 255): 257}
    for bad, frag in [(" # not a comment\n", "Invalid merge rule format"),   # '#' must be column 0
                      ("   \n", "Invalid merge rule format"),                  # blank but not empty
                      ("97 98 # c\n", "Invalid merge rule format"),           # inline comment
                      ("-1 5\n", "Failed to parse first byte value: invalid digit"),
                      ("5 300\n", "Failed to parse second byte value: number too large"),
                      ("5 +\n", "Failed to parse second byte value: invalid digit"),
                      ("1.0 2\n", "Failed to parse first byte value: invalid digit")]:
        with pytest.raises(oracle.OracleError) as ei:
            load(bad)
        assert ei.value.kind == oracle.ORA_INVALID_DATA and frag in ei.value.message, (bad, ei.value.message)
    with pytest.raises(oracle.OracleError) as ei:
        load(b"97 98\n\xff\xfe\n", binary=True)
    assert ei.value.kind == oracle.ORA_INVALID_DATA and "UTF-8" in ei.value.message
    # 65 280 rules is the last count whose ids fit u16; one more is rejected (DESIGN.md divergence)
    lines = "".join(f"{i % 256} {(i // 256) % 256}\n" for i in range(65280))
    assert max(load(lines).values()) == 65535
    with pytest.raises(oracle.OracleError):
        load(lines + "1 1\n")


# ---- derived vectors (frozen output of the independent Python model) ---------------------------------

@pytest.mark.parametrize("block", ["hand", "random"])
def test_derived_vectors(oracle, block):
    for row in DER[block]:
        md = merges_dict(row["merges"])
        data = bytes.fromhex(row["input_hex"])
        chunk = row["chunk"] or max(len(data), 1)
        got = oracle.run_buffer("bpe", data, chunk, 3, oracle.Merges(md))
        assert bytes(got) == be(row["tokens"]), row


# ---- property tests: oracle == python model == closed parallel form -----------------------------------

pairs = st.tuples(st.integers(0, 5), st.integers(0, 5))


@settings(max_examples=300, deadline=None)
@given(st.dictionaries(pairs, st.integers(0, 7), max_size=10), st.lists(st.integers(0, 3), max_size=80),
       st.integers(1, 40))
def test_oracle_equals_model_general_maps(oracle, merges, data, chunk):
    """General maps (chains, cycles, values that are also key components) over a tiny alphabet."""
    data = bytes(data)
    m = oracle.Merges(merges)
    assert bytes(oracle.run_buffer("bpe", data, chunk, 2, m)) == pm.run_buffer("bpe", data, chunk, merges)
    res, sweeps = oracle.process_chunk("bpe", data, m, want_sweeps=True)
    assert res == pm.bpe_chunk(data, merges)
    assert sweeps >= (1 if data else 0)


@settings(max_examples=300, deadline=None)
@given(st.dictionaries(st.tuples(st.integers(0, 3), st.integers(0, 3)), st.integers(256, 300), max_size=12),
       st.lists(st.integers(0, 3), max_size=120), st.integers(1, 50))
def test_parallel_form_equals_sequential_sweep(merges, data, chunk):
    """start[i] = m[i] & ~start[i-1] with m forced to 0 at chunk-last indices reproduces the
    reference's sequential sweep chunk by chunk, and for
    file-style tables (byte keys, ids >= 256) one sweep is already the fixpoint."""
    toks = list(data)
    ends = [i for i in range(len(toks)) if (i + 1) % chunk == 0]
    one = pm.sweep_parallel_form(toks, merges, ends)
    want = []
    for s in range(0, len(toks), chunk):
        want += pm.bpe_sweep(toks[s:s + chunk], merges)[0]
    assert one == want
    assert pm.to_be(one) == pm.run_buffer("bpe", bytes(data), chunk, merges)


def test_file_to_file_matches_buffer(oracle, tmp_path):
    rng = random.Random(7)
    data = bytes(rng.choice(b"ab c") for _ in range(100_000))
    merges = {(97, 98): 256, (98, 97): 257, (97, 97): 258, (32, 97): 259}
    m = oracle.Merges(merges)
    (tmp_path / "in.bin").write_bytes(data)
    for mode, mm in [("basic", None), ("bpe", m), ("passthrough", None)]:
        oracle.run_files(mode, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), 4096, 4, mm, 0xFF01)
        want = bytes(oracle.run_buffer(mode, data, 4096, 4, mm, 0xFF01))
        assert (tmp_path / "out.bin").read_bytes() == want
        assert want == pm.run_buffer(mode, data, 4096, merges, 0xFF01)
    (tmp_path / "empty").write_bytes(b"")
    oracle.run_files("bpe", str(tmp_path / "empty"), str(tmp_path / "out.bin"), 4096, 4, m)
    assert (tmp_path / "out.bin").read_bytes() == b""
    with pytest.raises(oracle.OracleError) as ei:
        oracle.run_files("basic", str(tmp_path / "nope"), str(tmp_path / "out.bin"), 4096, 1)
    assert ei.value.kind == oracle.ORA_NOT_FOUND


# ---- detokenizer (no reference counterpart: the two restatements pin each other and the round trip) ----

def test_detokenize_round_trip_on_reference_vectors(oracle):
    """tokenize -> detokenize gives the input back on every BPE / basic vector the reference holds."""
    for row in REF["bpe"]:
        if "merges_file" in row:
            pairs = pm.load_bpe_merges_text(row["merges_file"])
        else:
            pairs = merges_dict(row["merges"])
        invertible = all(a < 256 and b < 256 and v >= 256 for (a, b), v in pairs.items()) and \
            len(set(pairs.values())) == len(pairs)
        data = row["input"].encode()
        toks = oracle.process_chunk("bpe", data, oracle.Merges(pairs))
        if invertible:
            assert bytes(oracle.detokenize(toks, oracle.Merges(pairs))) == data, row["src"]
            assert pm.detokenize(toks, pairs) == data
        else:
            with pytest.raises(oracle.OracleError) as e:
                oracle.detokenize(toks, oracle.Merges(pairs))
            assert e.value.kind == -2
    for row in REF["basic"]:
        data = row["input"].encode() if isinstance(row["input"], str) else bytes(row["input"])
        assert bytes(oracle.detokenize(oracle.process_chunk("basic", data))) == data


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(st.integers(97, 101), st.integers(97, 101)), max_size=12, unique=True),
       st.binary(max_size=300), st.integers(1, 64), st.booleans())
def test_detokenize_oracle_equals_model(oracle, keys, data, chunk, with_ct):
    data = bytes(97 + b % 6 for b in data)
    pairs = {k: 300 + 7 * i for i, k in enumerate(keys)}              # non-contiguous ids on purpose
    om = oracle.Merges(pairs)
    toks = bytes(oracle.run_buffer("bpe", data, chunk, 2, om, content_type_token=0xFF01 if with_ct else None))
    assert bytes(oracle.detokenize(toks, om, with_ct)) == data
    assert pm.detokenize(toks, pairs, with_ct) == data


def test_detokenize_errors(oracle):
    om = oracle.Merges({(97, 98): 256})
    for bad, kind in ((b"\x00", -3), (b"\x01\x01", -3), (b"\xff\x01\x00a", -3)):   # odd, unknown id, prefix when none expected
        with pytest.raises(oracle.OracleError) as e:
            oracle.detokenize(bad, om)
        assert e.value.kind == kind
    with pytest.raises(oracle.OracleError):
        oracle.detokenize(b"\x00a", om, has_content_type=True)        # prefix expected, absent
    assert bytes(oracle.detokenize(b"\xff\x03\x01\x00\x00c", om, has_content_type=True)) == b"abc"
    for pairs in ({(97, 98): 256, (98, 97): 256}, {(300, 98): 400}, {(97, 98): 99}):
        with pytest.raises(oracle.OracleError) as e:
            oracle.detokenize(b"", oracle.Merges(pairs))
        assert e.value.kind == -2
