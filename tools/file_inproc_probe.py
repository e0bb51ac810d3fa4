#!/usr/bin/env python
"""blt_run_tokenizer called repeatedly from one process (the Python binding's situation), with the stage log on:
where the wall time of a warm call goes.   BLT_LOG=1 python tools/file_inproc_probe.py [--bytes N] [--gpus G]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blt_b200 import _native as nat, synth

ap = argparse.ArgumentParser()
ap.add_argument("--bytes", type=int, default=1 << 30)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--merges", type=int, default=32768)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
d = "/dev/shm"
inp, outp, mp = (os.path.join(d, f"blt_probe_{os.getpid()}.{e}") for e in ("in", "out", "merges.txt"))
data = synth.text(args.bytes, synth.SEED_CONFIG[3])
l, r = synth.merges_from_sample(data, args.merges)
synth.write_merges_file(mp, l, r)
data.tofile(inp)
try:
    for rep in range(args.reps):
        print(f"---- call {rep}", file=sys.stderr, flush=True)
        t0 = time.perf_counter()
        nat.run_tokenizer(inp, outp, merges_file=mp, chunk_size="16MB", num_gpus=args.gpus)
        dt = time.perf_counter() - t0
        print(f"call {rep}: {dt:.3f} s = {args.bytes / dt / 1e9:.2f} GB/s, output {os.path.getsize(outp)} bytes", flush=True)
finally:
    for f in (inp, outp, mp):
        if os.path.exists(f):
            os.unlink(f)
