#!/usr/bin/env python
"""Small parity cases for compute-sanitizer runs (memcheck / racecheck): dense, sparse, walls, ragged
tails, general maps, basic.  Exits non-zero on any mismatch with the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from blt_b200 import _native as nat, synth
from oracle import oracle_ffi as ora

ctx = nat.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600_001
text = synth.text(n, 123)
cases = []
l, r = synth.merges_from_sample(text, 32768)
cases.append(("dense", text, {(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))}))
l, r = synth.merges_from_sample(text, 64)
cases.append(("sparse", text, {(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))}))
cases.append(("adversarial", synth.adversarial(n, 5), {p: 256 + i for i, p in enumerate(synth.adversarial_pairs())}))
cases.append(("chain", np.frombuffer((b"abcd" * (n // 4 + 1))[:n // 8], dtype=np.uint8), {(97, 98): 256, (256, 99): 257, (257, 100): 258, (258, 258): 97}))
bad = 0
for name, data, pairs in cases:
    s = ctx.bpe_from_pairs(pairs)
    om = ora.Merges(pairs)
    for chunk in (65536, 100001, data.size):
        got = s.tokenize_host(data, chunk_size=chunk)
        want = ora.run_buffer("bpe", data, chunk, 4, om)
        ok = np.array_equal(got, want)
        print(name, chunk, "ok" if ok else "MISMATCH", flush=True)
        bad += not ok
    s.close()
b = ctx.basic()
ok = np.array_equal(b.tokenize_host(text, chunk_size=70000), ora.run_buffer("basic", text, 70000, 2))
print("basic", "ok" if ok else "MISMATCH")
bad += not ok
sys.exit(1 if bad else 0)
