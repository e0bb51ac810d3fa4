#!/usr/bin/env python
"""Randomised parity fuzzing on the GPU: random sizes, chunk sizes, tables (dense, sparse, holes), all sweep
variants and dense-pass settings, every entry point (resident, host pipeline, per-chunk, detokenizer round
trip), each output compared with the CPU oracle.  python tools/fuzz_gpu.py --seconds 200 [--seed S]"""
import argparse, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from blt_b200 import _native as nat
from oracle import oracle_ffi as ora

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=120)
ap.add_argument("--seed", type=int, default=int(time.time()))
ap.add_argument("--only", type=int, default=-1, help="replay one case index of this seed, verbosely")
args = ap.parse_args()
print("seed", args.seed, flush=True)
torch.cuda.set_device(0)
stream = torch.cuda.current_stream().cuda_stream
t_end = time.time() + args.seconds
cases = fails = 0
index = args.only if args.only >= 0 else 0
while time.time() < t_end:
    rng = random.Random(args.seed * 1000003 + index)
    nrng = np.random.default_rng(args.seed * 1000003 + index)
    os.environ["BLT_SWEEP_VARIANT"] = str(rng.choice([0, 1, 2]))
    os.environ["BLT_DENSE"] = rng.choice(["0", "1", "always"])
    ctx = nat.Context(0)
    alpha = rng.choice([2, 3, 5, 26, 256])
    syms = rng.sample(range(256), alpha)
    n = rng.choice([rng.randint(0, 70), rng.randint(70, 5000), rng.randint(5000, 300000), rng.randint(300000, 6 << 20)])
    style = rng.choice(["iid", "runs", "zipf"])
    if style == "iid":
        data = nrng.choice(np.array(syms, dtype=np.uint8), size=n)
    elif style == "runs":
        lens = nrng.geometric(rng.choice([0.5, 0.1, 0.01, 0.0005]), size=max(1, n // 2 + 1))
        vals = nrng.choice(np.array(syms, dtype=np.uint8), size=lens.size)
        data = np.repeat(vals, lens)[:n].astype(np.uint8)
        if data.size < n:
            data = np.resize(data, n)
    else:
        p = 1.0 / np.arange(1, alpha + 1); p /= p.sum()
        data = nrng.choice(np.array(syms, dtype=np.uint8), size=n, p=p)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    density = rng.choice([1.0, 0.95, 0.6, 0.2, 0.02])
    allp = [(a, b) for a in syms for b in syms] if alpha <= 26 else [(rng.choice(syms), rng.choice(syms)) for _ in range(20000)]
    rng.shuffle(allp)
    keys = list(dict.fromkeys(allp))[: max(0, int(len(allp) * density))]
    gap = rng.choice([1, 1, 1, 3])
    pairs = {k: 256 + gap * i for i, k in enumerate(keys) if 256 + gap * i < 65536}
    om = ora.Merges(pairs)
    s = ctx.bpe_from_pairs(pairs)
    chunk = rng.choice([0, 2, 16, 100, 4096, 4098, 65536, 100001, 1 << 20])
    eff = chunk if chunk and chunk < n else max(n, 1)
    want = ora.run_buffer("bpe", data, eff, 4, om)
    ok = True
    bad = []
    # resident
    d_in = torch.from_numpy(data).cuda() if n else torch.empty(16, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(2 * n + 16, dtype=torch.uint8, device="cuda")
    nc = max(1, (n + eff - 1) // eff)
    d_ends = torch.zeros(nc, dtype=torch.int64, device="cuda")
    for rep in range(rng.choice([1, 3])):
        got_len = s.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), 2 * n, d_ends.data_ptr(), stream)
        got = d_out[:got_len].cpu().numpy()
        if not np.array_equal(got, want): bad.append(f"resident rep {rep}: len {got_len} vs {want.size}")
        if n and int(d_ends[-1].item()) != want.size: bad.append(f"chunk_ends[-1] {int(d_ends[-1].item())} vs {want.size}")
    # host pipeline (pageable -> staged) and per-chunk call
    if not np.array_equal(s.tokenize_host(data, chunk_size=chunk or max(n, 1)), want): bad.append("tokenize_host")
    if n and n <= (1 << 20):
        if not np.array_equal(s.process_chunk(data), np.frombuffer(ora.process_chunk("bpe", data, om), dtype=np.uint8)): bad.append("process_chunk")
    # detokenizer round trip
    try:
        if not np.array_equal(s.detokenize_host(want), data): bad.append("detokenize round trip")
    except nat.BltError as e:
        bad.append(f"detokenize raised {e}")
    cases += 1
    ok = not bad
    if not ok:
        fails += 1
        print("MISMATCH", index, bad, dict(variant=os.environ["BLT_SWEEP_VARIANT"], dense=os.environ["BLT_DENSE"], n=n, chunk=chunk, alpha=alpha,
                               style=style, density=density, rules=len(pairs), gap=gap), flush=True)
    s.close()
    ctx.close()
    if args.only >= 0:
        break
    index += 1
print(f"fuzz: {cases} cases, {fails} mismatches, seed {args.seed}")
sys.exit(1 if fails else 0)
