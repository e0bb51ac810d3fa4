#!/usr/bin/env python
"""The drop-in as the reference would use it: T host threads, each calling
blt_process_chunk on 16 MiB slices of ONE pageable input buffer (the reference's mmap) into its own pageable
output buffer (the reference's Vec<u8>), results consumed in chunk order.  Prints aggregate input GB/s."""
import argparse, json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from blt_b200 import _native as nat, synth

ap = argparse.ArgumentParser()
ap.add_argument("--bytes", type=int, default=1 << 30)
ap.add_argument("--threads", default="1,2,4,8,16")
ap.add_argument("--merges", type=int, default=32768)
args = ap.parse_args()
n, chunk = args.bytes, 16 << 20
data = synth.text(n, synth.SEED_CONFIG[3])
l, r = synth.merges_from_sample(data, args.merges)
ctx = nat.Context(0)
strat = ctx.bpe_from_pairs({(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))})
n_chunks = (n + chunk - 1) // chunk
want = strat.tokenize_host(data, chunk_size=chunk)

for T in [int(t) for t in args.threads.split(",")]:
    best = None
    for rep in range(3):
        outs = [None] * n_chunks
        nxt = [0]
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    k = nxt[0]
                    nxt[0] += 1
                if k >= n_chunks:
                    return
                buf = np.empty(2 * chunk, dtype=np.uint8)          # a fresh "Vec<u8>" per chunk, like the reference
                outs[k] = strat.process_chunk(data[k * chunk:(k + 1) * chunk], out=buf)

        t0 = time.perf_counter()
        ths = [threading.Thread(target=worker) for _ in range(T)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    got = np.concatenate(outs)
    print(json.dumps({"api": "blt_process_chunk, pageable buffers", "threads": T, "bytes": n, "seconds": round(best, 4),
                      "input_GBps": round(n / best / 1e9, 2), "matches_pipeline_output": bool(np.array_equal(got, want))}), flush=True)
