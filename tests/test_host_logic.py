"""CPU-only checks of the product's host side (libblt_cuda.so without a device): the C ABI exports
every symbol the header declares, the merges loader / size grammar / chunk sizing / thread count
reproduce the reference's golden vectors, the Python surface mirrors blt_python's, and every compute
entry point fails loudly without a CUDA device (no CPU fallback)."""
import ctypes
import json
import os
import re
import subprocess
import sys

import pytest

import blt_b200
from blt_b200 import _native as nat

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))
BLT = os.path.join(ROOT, "blt_b200", "lib", "blt")


def has_gpu() -> bool:
    try:
        return nat.device_count() > 0
    except nat.BltError:
        return False


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "blt_cuda.h")).read()
    declared = set(re.findall(r"BLT_API\s+[\w\s\*]+?\b(blt_\w+)\s*\(", hdr))
    assert len(declared) >= 20, declared
    lib = ctypes.CDLL(nat.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", nat.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (blt_\w+)", out))
    assert declared <= exported
    assert exported - declared == set(), "exported but undeclared: %s" % (exported - declared)


def test_version_matches_reference_package_version():
    assert blt_b200.version() == "0.2.2" == blt_b200.__version__ == nat.version()


@pytest.mark.parametrize("row", REF["merges_files"], ids=lambda r: r["src"])
def test_product_loader_golden(row, tmp_path):
    p = tmp_path / "m.txt"
    p.write_text(row["text"])
    want = {(a, b): v for a, b, v in row["map"]}
    assert nat.load_bpe_merges(str(p)) == want
    assert blt_b200.load_bpe_merges(str(p)) == want


@pytest.mark.parametrize("row", REF["merges_file_errors"], ids=lambda r: r["src"])
def test_product_loader_errors(row, tmp_path):
    kinds = {"InvalidData": nat.ERR_INVALID_DATA, "NotFound": nat.ERR_NOT_FOUND}
    p = tmp_path / "m.txt"
    if not row.get("missing_file"):
        p.write_text(row["text"])
    with pytest.raises(nat.BltError) as ei:
        nat.load_bpe_merges(str(p))
    assert ei.value.code == kinds[row["kind"]]
    if "contains" in row:
        assert row["contains"] in ei.value.message
    # Python surface: IOError for a missing file (blt_python/tests/test_tokenizer.py:229-232)
    with pytest.raises((IOError, ValueError)):
        blt_b200.load_bpe_merges(str(p))


def test_product_loader_matches_oracle_on_corner_cases(oracle, tmp_path):
    cases = [b"97 98", b"97 98\r\n99 100\r\n", b"  97\t 98  \n", b"+97 098\n", b"\n\n#x\n0 0\n255 255\n",
             b" # not a comment\n", b"   \n", b"97 98 # c\n", b"-1 5\n", b"5 300\n", b"5 +\n", b"1.0 2\n",
             b"97 98\n\xff\xfe\n", b"1 2\n1 2\n1 2\n", b"97\xc2\xa098\n", b"97 98\r", b"#\n", b"256 1\n", b"1 2 3\n",
             b"9999999999999999999999 1\n", b"12a 1\n", b"1 \xe2\x80\x8398\n"]
    for i, raw in enumerate(cases):
        p = tmp_path / f"c{i}.txt"
        p.write_bytes(raw)
        try:
            want = ("ok", oracle.Merges.from_file(str(p)).to_dict())
        except oracle.OracleError as e:
            want = ("err", e.kind, e.message)
        try:
            got = ("ok", nat.load_bpe_merges(str(p)))
        except nat.BltError as e:
            got = ("err", e.code, e.message)
        assert got == want, raw
    lines = "".join(f"{i % 256} {(i // 256) % 256}\n" for i in range(65280))
    (tmp_path / "big.txt").write_text(lines)
    assert max(nat.load_bpe_merges(str(tmp_path / "big.txt")).values()) == 65535
    (tmp_path / "big.txt").write_text(lines + "1 1\n")
    with pytest.raises(nat.BltError):
        nat.load_bpe_merges(str(tmp_path / "big.txt"))


def test_chunk_size_grammar_and_clamps(oracle):
    for s, v in REF["chunk_size_parse"]["valid"]:
        assert nat.parse_chunk_size(s) == v
    for s in REF["chunk_size_parse"]["invalid"] + ["+", "1 KB", "99999999999999999999", "18014398509481984KB"]:
        with pytest.raises(nat.BltError) as ei:
            nat.parse_chunk_size(s)
        assert ei.value.code == nat.ERR_INVALID_INPUT
        with pytest.raises(oracle.OracleError):
            oracle.parse_chunk_size(s)
    for s in ["+5KB", "007", "0", "16mb", "\t3Mb\n", "1kB "]:
        assert nat.parse_chunk_size(s) == oracle.parse_chunk_size(s)
    for cli, want in REF["chunk_size_clamp"]["cases"]:
        assert nat.effective_chunk_size(cli, 4, 80, 64 << 30) == want
    for ram in [1 << 28, 1 << 30, 8 << 30, 64 << 30, 2 << 40]:
        for threads in [1, 4, 128, 1000]:
            for cap in [0, 1, 50, 80, 100]:
                assert nat.effective_chunk_size(None, threads, cap, ram) == oracle.effective_chunk_size(None, threads, cap, ram)
    assert (1 << 20) <= nat.effective_chunk_size(None, 4, 80, 0) <= (16 << 20)   # probes this host


def test_thread_count_and_content_type_tokens():
    for v, want in REF["thread_count"]["cases"]:
        assert nat.determine_thread_count(v) == want
    assert nat.determine_thread_count(None) == len(os.sched_getaffinity(0))
    t = REF["content_type_tokens"]
    assert [nat.content_type_token(i) for i in range(4)] == [t["text"], t["audio"], t["bin"], t["video"]]
    assert nat.content_type_token(-1) == 0


def test_shard_chunks_is_floor_k_g_over_k():
    for n_chunks in [1, 2, 7, 64, 65, 512, 513]:
        for g in [1, 2, 3, 4, 8]:
            b = nat.shard_chunks(n_chunks, g)
            assert b[0] == 0 and b[-1] == n_chunks and all(x <= y for x, y in zip(b, b[1:]))
            for k in range(n_chunks):
                owner = (k * g) // n_chunks
                assert b[owner] <= k < b[owner + 1]


def test_file_pipeline_deals_chunks_round_robin():
    """blt_run_tokenizer's sharding (include/blt_cuda.h): chunk k -> GPU k mod G; every GPU gets floor or ceil of K/G chunks."""
    for g in [1, 2, 3, 4, 8]:
        owners = [nat.file_chunk_device(k, g) for k in range(67)]
        assert owners == [k % g for k in range(67)]
        counts = [owners.count(d) for d in range(g)]
        assert max(counts) - min(counts) <= 1
    assert nat.file_chunk_device(5, 0) == 0


# ---- Python surface, mirroring blt_python/tests/test_tokenizer.py (the rows that need no device) ----

def test_python_surface_shape():
    for name in ["ByteTokenizer", "load_bpe_merges", "version", "__version__"]:
        assert hasattr(blt_b200, name)
    t = blt_b200.ByteTokenizer()
    assert "ByteTokenizer" in str(t)
    assert repr(t) == "ByteTokenizer(merges=0, content_type=None, threads=None, chunk_size=None, memory_cap=None)"
    assert "merges=2" in str(blt_b200.ByteTokenizer(merges={(97, 98): 256, (99, 100): 257}))
    assert 'content_type=Some("Text")' in str(blt_b200.ByteTokenizer(content_type="Text"))
    assert 'content_type=Some("Bin")' in str(blt_b200.ByteTokenizer(content_type="Bin"))
    assert repr(blt_b200.ByteTokenizer(threads=2, chunk_size="1MB", memory_cap=50)).endswith(
        'threads=Some(2), chunk_size=Some("1MB"), memory_cap=Some(50))')
    with pytest.raises(ValueError):
        blt_b200.ByteTokenizer(content_type="Invalid")
    with pytest.raises(ValueError):
        blt_b200.ByteTokenizer(memory_cap=150)
    with pytest.raises(IOError):
        blt_b200.load_bpe_merges("non_existent_file.txt")


def test_blt_alias_package_and_lazy_library():
    """`import blt` is the reference's module name (blt_python/python/blt/__init__.py:12-16); importing the package or
    the workload generators must not map libblt_cuda.so (bench.py's reference arm relies on it)."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import blt, blt_b200\n"
            "from blt_b200 import synth\n"
            "assert blt.ByteTokenizer is blt_b200.ByteTokenizer and blt.load_bpe_merges is blt_b200.load_bpe_merges\n"
            "assert sorted(blt.__all__) == ['ByteTokenizer', '__version__', 'load_bpe_merges', 'version']\n"
            "d = synth.mixed(1 << 16)\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libblt_cuda' not in maps, 'importing the package loaded the CUDA library'\n"
            "assert blt.__version__ == blt.version() == '0.2.2'\n"
            "assert 'libblt_cuda' in open('/proc/self/maps').read()\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_rust_shim_binds_only_declared_symbols():
    """integration/rust/*.rs (untested: no Rust toolchain here) may only bind functions the header declares and the
    library exports, with the header's arity."""
    import re
    hdr = open(os.path.join(ROOT, "include", "blt_cuda.h")).read()
    lib = nat.lib()
    for f in ("cuda_strategy.rs", "select_strategy.rs"):
        src = open(os.path.join(ROOT, "integration", "rust", f)).read()
        for name, args in re.findall(r"fn (blt_\w+)\(([^)]*)\)", src):
            assert hasattr(lib, name), name
            m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, hdr, re.S)
            assert m, name
            n_rust = 0 if not args.strip() else len([a for a in args.split(",") if a.strip()])
            c_args = m.group(1).strip()
            n_c = 0 if c_args in ("", "void") else len(c_args.split(","))
            assert n_rust == n_c, (name, n_rust, n_c)


# ---- no CPU fallback --------------------------------------------------------------------------------

@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a CUDA device")
def test_compute_fails_loudly_without_a_device(tmp_path):
    with pytest.raises(nat.BltError) as ei:
        nat.Context(0)
    assert ei.value.code == nat.ERR_NO_DEVICE
    (tmp_path / "in.bin").write_bytes(b"hello world")
    with pytest.raises(RuntimeError):
        blt_b200.ByteTokenizer().tokenize_file(str(tmp_path / "in.bin"), str(tmp_path / "out.bin"))
    r = subprocess.run([BLT, "-i", str(tmp_path / "in.bin"), "-o", str(tmp_path / "o2.bin")], capture_output=True)
    assert r.returncode == 1 and b"Error running tokenizer" in r.stderr


# ---- CLI surface that needs no device (src/main.rs:8-106) ---------------------------------------------

def test_cli_flags_and_exit_codes(tmp_path):
    r = subprocess.run([BLT, "--version"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "blt 0.2.2"
    r = subprocess.run([BLT, "--help"], capture_output=True, text=True)
    for flag in ["-i, --input", "-o, --output", "--merges", "--passthrough", "--type", "--threads", "--memcap",
                 "--chunksize", "-h, --help", "-V, --version"]:
        assert flag in r.stdout
    # passthrough is a copy: tests/cli.rs:196-214 (stdin -> stdout), and file -> file
    r = subprocess.run([BLT, "--passthrough"], input=b"passthrough test", capture_output=True)
    assert r.returncode == 0 and r.stdout == b"passthrough test"
    r = subprocess.run([BLT, "--passthrough", "--type", "video"], input=b"x", capture_output=True)
    assert r.stdout == b"\xff\x04x"
    (tmp_path / "in.bin").write_bytes(b"abc" * 1000)
    r = subprocess.run([BLT, "--passthrough", "-i", str(tmp_path / "in.bin"), "-o", str(tmp_path / "out.bin")])
    assert r.returncode == 0 and (tmp_path / "out.bin").read_bytes() == b"abc" * 1000
    # clap usage errors -> exit 2
    for bad in [["--type", "TEXT"], ["--threads", "x"], ["--memcap", "256"], ["--bogus"], ["--merges"], ["-m", "x"]]:
        assert subprocess.run([BLT] + bad, capture_output=True, input=b"").returncode == 2, bad
    # CoreConfig::new_from_cli errors -> `Error: ...`, exit 1, before any IO happens
    out = tmp_path / "never.bin"
    r = subprocess.run([BLT, "--chunksize", "1gb", "-o", str(out)], capture_output=True, input=b"x")
    assert r.returncode == 1 and b"InvalidInput" in r.stderr and not out.exists()
    (tmp_path / "bad.txt").write_text("97 abc\n")
    r = subprocess.run([BLT, "--merges", str(tmp_path / "bad.txt"), "-o", str(out)], capture_output=True, input=b"x")
    assert r.returncode == 1 and b"Failed to load BPE merges: Failed to parse second byte value" in r.stderr
    assert not out.exists()
    r = subprocess.run([BLT, "--passthrough", "-i", str(tmp_path / "missing")], capture_output=True)
    assert r.returncode == 1 and b"Error running tokenizer" in r.stderr


def test_select_merges_host_rule():
    """blt_select_merges (host only): most frequent first, ties by b0*256+b1 ascending, optional padding."""
    import numpy as np
    from blt_b200 import _native as nat
    counts = np.zeros(65536, dtype=np.uint64)
    counts[(101 << 8) | 32] = 50      # "e "
    counts[(116 << 8) | 104] = 50     # "th"  (tie: 101*256+32 < 116*256+104)
    counts[(0 << 8) | 1] = 7
    counts[(255 << 8) | 255] = 900
    l, r = nat.select_merges(counts, 3)
    assert list(zip(l.tolist(), r.tolist())) == [(255, 255), (101, 32), (116, 104)]
    l, r = nat.select_merges(counts, 10)                     # only 4 pairs occur
    assert len(l) == 4 and (l[3], r[3]) == (0, 1)
    l, r = nat.select_merges(counts, 7, pad_unobserved=True)  # padded with (0,0), (0,2), (0,3): (0,1) occurs
    assert list(zip(l.tolist(), r.tolist()))[4:] == [(0, 0), (0, 2), (0, 3)]
    with pytest.raises(nat.BltError):
        nat.select_merges(counts, 65281)


def test_header_is_plain_c_and_links(tmp_path):
    """include/blt_cuda.h must compile as C (the boundary is a C ABI: cgo / Rust FFI / ctypes bind to it) and a
    C program must link against libblt_cuda.so and use the host-only entry points."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "blt_b200", "lib")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include "blt_cuda.h"
#include <stdio.h>
#include <string.h>
int main(void) {
    size_t v = 0;
    if (blt_parse_chunk_size("16MB", &v) != BLT_OK || v != 16u * 1024u * 1024u) return 1;
    if (blt_parse_chunk_size("1gb", &v) != BLT_ERR_INVALID_INPUT) return 2;
    if (blt_content_type_token(BLT_CONTENT_TEXT) != 0xFF01) return 3;
    if (blt_effective_chunk_size(1, 1, 4, 80, 0) != 256u * 1024u) return 4;   /* clamp up, chunking.rs:29 */
    if (strlen(blt_version()) == 0) return 5;
    blt_core_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.chunk_size = "bogus";
    if (blt_run_tokenizer(&cfg) != BLT_ERR_INVALID_INPUT) return 6;          /* config errors come before any device use */
    printf("ok %s\n", blt_version());
    return 0;
}
''')
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lblt_cuda", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok "), (r.returncode, r.stdout, r.stderr)


def test_bench_reference_arm_is_clean_and_prints_the_same_config():
    """bench.py --impl reference: runs on the oracle only (no libblt_cuda.so mapped), prints the contract's keys and
    the same `config` object the product arm builds (reference-arm hygiene)."""
    import importlib.util
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--bytes", str(32 << 20), "--merges", "2048"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["value"] > 0
    assert line["notes"]["maps_libblt_cuda"] is False
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class A:
        bytes, merges = 32 << 20, 2048
    assert line["config"] == mod.workload_config(A)
