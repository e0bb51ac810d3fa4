#!/usr/bin/env python
"""Replays given byte inputs through blt_process_resident inside canary-guarded device buffers and reports
which guard (or the input itself) a launch damaged.  Debug helper for tools/fuzz_gpu.py findings."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from blt_b200 import _native as nat
from oracle import oracle_ffi as ora

def run(name, data, pairs, chunk, variant, dense, reps=3):
    os.environ["BLT_SWEEP_VARIANT"] = str(variant); os.environ["BLT_DENSE"] = dense
    ctx = nat.Context(0); s = ctx.bpe_from_pairs(pairs); om = ora.Merges(pairs)
    n = data.size
    eff = chunk if chunk and chunk < n else max(n, 1)
    want = ora.run_buffer("bpe", data, eff, 2, om)
    G = 4096
    nc = max(1, (n + eff - 1) // eff)
    al = lambda x: (x + 255) // 256 * 256
    o_in = G; o_out = o_in + al(n) + G; o_ends = o_out + al(2 * n + 16) + G; total = o_ends + al(8 * nc) + G
    buf = torch.full((total,), 0xA5, dtype=torch.uint8, device="cuda")
    buf[o_in:o_in + n] = torch.from_numpy(data).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    base = buf.data_ptr()
    for rep in range(reps):
        ln = s.process_resident(base + o_in, n, chunk, base + o_out, 2 * n, base + o_ends, stream)
        h = buf.cpu().numpy()
        issues = []
        if not np.array_equal(h[o_in:o_in + n], data): issues.append("INPUT MODIFIED at %s" % np.nonzero(h[o_in:o_in + n] != data)[0][:8])
        for nm, lo, hi in (("guard0", 0, o_in), ("guard after in", o_in + n, o_out), ("guard after out(cap 2n)", o_out + 2 * n, o_ends), ("guard after ends", o_ends + 8 * nc, total)):
            g = h[lo:hi]
            if not np.all(g == 0xA5): issues.append(f"{nm} damaged at +{np.nonzero(g != 0xA5)[0][:8]}")
        got = h[o_out:o_out + ln]
        if not np.array_equal(got, want):
            k = int(np.argmax(got[:min(len(got), len(want))] != want[:min(len(got), len(want))])) if len(got) and len(want) else -1
            issues.append(f"output differs: len {ln} vs {want.size}, first diff at byte {k}")
        ends = h[o_ends:o_ends + 8 * nc].view(np.int64)
        if n and ends[-1] != want.size: issues.append(f"ends[-1]={ends[-1]}")
        print(name, "rep", rep, "OK" if not issues else issues, flush=True)
    s.close(); ctx.close()

rng = np.random.default_rng(1)
# tiny inputs, repeated calls
syms = np.array([97, 98, 99], dtype=np.uint8)
pairs = {(97, 97): 256, (97, 98): 257, (98, 97): 258, (98, 98): 259, (99, 97): 260, (97, 99): 261, (99, 99): 262, (98, 99): 263}
for n in (6, 15, 21, 35):
    data = rng.choice(syms, size=n)
    for variant, dense in ((1, "0"), (2, "always"), (0, "always"), (0, "0")):
        run(f"n={n} v={variant} dense={dense} chunk=0", data, pairs, 0, variant, dense)
# tiny chunks
p5 = {(a, b): 256 + i for i, (a, b) in enumerate([(x, y) for x in (97, 98, 99, 100, 101) for y in (97, 98, 99, 100, 101)][:23])}
data = rng.choice(np.array([97, 98, 99, 100, 101], dtype=np.uint8), size=253)
for variant in (0, 1, 2):
    for chunk in (2, 16):
        run(f"n=253 v={variant} chunk={chunk}", data, p5, chunk, variant, "0", reps=1)
data = rng.choice(np.array([97, 98, 99], dtype=np.uint8), size=2119)
run("n=2119 chunk=16 v=0 always", data, {(97, 98): 256}, 16, 0, "always", reps=2)
