"""Generates tests/golden/derived_vectors.json from oracle/py_model.py.

These are NOT reference-held vectors (the reference's tests pin none of this behaviour): they are
answers derived from blt_core/src/tokenizer.rs:63-86 and blt_core/src/pipeline.rs:73-81 by the
independent Python model, frozen so that the C++ oracle and the CUDA path are both checked against a
fixed file.  The first block carries the hand-derived expectations of SURVEY.md section 4, which the
script asserts before writing.  Run:  python tests/golden/make_derived_vectors.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import py_model as pm  # noqa: E402

A, B, C = 97, 98, 99
HAND = [  # (merges, input, chunk_size or None, expected tokens)  -- SURVEY.md section 4
    ({(A, A): 256}, b"aaaaa", None, [256, 256, 97]),
    ({(B, A): 256, (A, A): 257}, b"baaa", None, [256, 257]),
    ({(A, B): 256, (B, A): 257}, b"abab", None, [256, 256]),
    ({(A, A): 256, (A, B): 257}, b"aab", None, [256, 98]),
    ({(B, C): 256, (A, B): 257}, b"abc", None, [257, 99]),
    ({(A, B): 99, (C, B): 100}, b"abb", None, [100]),
    ({(A, A): 97}, b"aaaaaaaa", None, [97]),
    ({(A, B): 256}, b"abab", 3, [256, 97, 98]),
    ({(A, A): 256}, b"aaaaaa", 3, [256, 97, 256, 97]),
]


def tokens_of(be: bytes):
    return [int.from_bytes(be[i:i + 2], "big") for i in range(0, len(be), 2)]


def main():
    out = {"_comment": __doc__.strip(), "hand": [], "random": []}
    for merges, data, chunk, expect in HAND:
        got = tokens_of(pm.run_buffer("bpe", data, chunk or max(len(data), 1), merges))
        assert got == expect, (merges, data, chunk, got, expect)
        out["hand"].append({"merges": [[a, b, v] for (a, b), v in merges.items()],
                            "input_hex": data.hex(), "chunk": chunk, "tokens": expect})
    rng = random.Random(0xB17)
    for case in range(200):
        alpha = [rng.randrange(256) for _ in range(rng.choice([2, 3, 4, 8]))]
        n_rules = rng.randrange(0, 12)
        merges = {}
        general = case % 4 == 3  # chains / cycles / values < 256 through the pair API
        for i in range(n_rules):
            pool = alpha + ([256 + j for j in range(4)] if general else [])
            val = rng.choice(pool) if (general and rng.random() < 0.4) else 256 + i
            merges[(rng.choice(pool), rng.choice(pool))] = val
        n = rng.choice([0, 1, 2, 3, 5, 16, 17, 31, 32, 33, 64, 100, 257])
        data = bytes(rng.choice(alpha) for _ in range(n))
        chunk = rng.choice([None, 1, 2, 3, 7, 16, 32, 33])
        toks = tokens_of(pm.run_buffer("bpe", data, chunk or max(n, 1), merges))
        out["random"].append({"merges": [[a, b, v] for (a, b), v in merges.items()],
                              "input_hex": data.hex(), "chunk": chunk, "tokens": toks})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "derived_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, len(out["hand"]), len(out["random"]))


if __name__ == "__main__":
    main()
