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
    v = rng.choice([0, 1, 2, 3, 3, 4, 4, None, None])   # None: the library picks the exact form per call
    if v is None:
        os.environ.pop("BLT_SWEEP_VARIANT", None)
    else:
        os.environ["BLT_SWEEP_VARIANT"] = str(v)
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
    density = rng.choice([1.0, 1.0, 0.95, 0.6, 0.2, 0.02])
    if density == 1.0 and alpha < 256 and n and rng.random() < 0.6:
        # a table that covers every pair of the alphabet and a few foreign bytes: the dense pass holds up to the first
        # chunk with one of them, the exact sweep redoes the rest
        foreign = next(b for b in range(256) if b not in syms)
        for _ in range(rng.choice([1, 1, 2, 5])):
            data[rng.randrange(n)] = foreign
    allp = [(a, b) for a in syms for b in syms] if alpha <= 26 else [(rng.choice(syms), rng.choice(syms)) for _ in range(20000)]
    rng.shuffle(allp)
    keys = list(dict.fromkeys(allp))[: max(0, int(len(allp) * density))]
    gap = rng.choice([1, 1, 1, 3])
    pairs = {k: 256 + gap * i for i, k in enumerate(keys) if 256 + gap * i < 65536}
    om = ora.Merges(pairs)
    s = ctx.bpe_from_pairs(pairs)
    chunk = rng.choice([0, 0, 2, 16, 100, 4096, 4098, 16384, 32768, 65536, 100001, 1 << 20])
    eff = chunk if chunk and chunk < n else max(n, 1)
    want = ora.run_buffer("bpe", data, eff, 4, om)
    ok = True
    bad = []
    # resident, inside canary-guarded buffers: output capacity exactly 2n, chunk_ends exactly one entry per chunk
    G = 1024
    al = lambda x: (x + 255) // 256 * 256
    nc = max(1, (n + eff - 1) // eff)
    o_in = G; o_out = o_in + al(n) + G; o_ends = o_out + al(2 * n) + G; total = o_ends + al(8 * nc) + G
    buf = torch.full((total,), 0xA5, dtype=torch.uint8, device="cuda")
    if n: buf[o_in:o_in + n] = torch.from_numpy(data).cuda()
    base = buf.data_ptr()
    for rep in range(rng.choice([1, 3])):
        got_len = s.process_resident(base + o_in, n, chunk, base + o_out, 2 * n, base + o_ends, stream)
        h = buf.cpu().numpy()
        if not np.array_equal(h[o_in:o_in + n], data): bad.append(f"resident rep {rep}: input modified")
        for lo, hi in ((0, o_in), (o_in + n, o_out), (o_out + 2 * n, o_ends), (o_ends + 8 * nc, total)):
            if not np.all(h[lo:hi] == 0xA5): bad.append(f"resident rep {rep}: canary at {lo} damaged")
        if not np.array_equal(h[o_out:o_out + got_len], want): bad.append(f"resident rep {rep}: len {got_len} vs {want.size}")
        if n and int(h[o_ends:o_ends + 8 * nc].view(np.int64)[-1]) != want.size: bad.append("chunk_ends[-1]")
    # host pipeline (pageable -> staged) and per-chunk call
    if not np.array_equal(s.tokenize_host(data, chunk_size=chunk or max(n, 1)), want): bad.append("tokenize_host")
    if n and n <= (1 << 20):
        if not np.array_equal(s.process_chunk(data), np.frombuffer(ora.process_chunk("bpe", data, om), dtype=np.uint8)): bad.append("process_chunk")
    # detokenizer round trip
    try:
        if not np.array_equal(s.detokenize_host(want), data): bad.append("detokenize round trip")
    except nat.BltError as e:
        bad.append(f"detokenize raised {e}")
    # detokenizer on an arbitrary valid token stream (not a tokenizer output), capacity exactly the answer's size
    if pairs:
        idsv = np.array(sorted(set(pairs.values())), dtype=np.int64)
        nt = rng.choice([0, 1, 7, 255, 256, 257, rng.randint(0, 200000)])
        wide = nrng.random(nt) < rng.choice([0.0, 0.3, 1.0])
        toks = np.where(wide, idsv[nrng.integers(0, len(idsv), nt)], nrng.integers(0, 256, nt)).astype(">u2").view(np.uint8)
        if len(set(pairs.values())) == len(pairs):
            dw = ora.detokenize(toks, om)
            o_t = G; o_b = o_t + al(toks.size) + G; tot2 = o_b + al(dw.size) + G
            b2 = torch.full((tot2,), 0x5A, dtype=torch.uint8, device="cuda")
            if toks.size: b2[o_t:o_t + toks.size] = torch.from_numpy(np.ascontiguousarray(toks)).cuda()
            ln = s.detokenize_resident(b2.data_ptr() + o_t, toks.size, b2.data_ptr() + o_b, dw.size, stream)
            h2 = b2.cpu().numpy()
            if ln != dw.size or not np.array_equal(h2[o_b:o_b + ln], dw): bad.append("detokenize_resident")
            for lo, hi in ((0, o_t), (o_t + toks.size, o_b), (o_b + dw.size, tot2)):
                if not np.all(h2[lo:hi] == 0x5A): bad.append(f"detokenize canary at {lo}")
    # basic strategy and the pair histogram on the same bytes
    b = ctx.basic()
    if not np.array_equal(b.tokenize_host(data, chunk_size=chunk or max(n, 1)), ora.run_buffer("basic", data, eff, 2)): bad.append("basic")
    b.close()
    if n >= 2:
        keys2 = (data[:-1].astype(np.uint32) << 8) | data[1:]
        if not np.array_equal(ctx.count_pairs(data), np.bincount(keys2, minlength=65536).astype(np.uint64)): bad.append("pair histogram")
    # a general map (second-level rules on produced ids): the multi-sweep hash path
    if n <= 200000 and pairs:
        gp = dict(pairs)
        ids = list(pairs.values())
        for i in range(min(8, len(ids))):
            gp[(ids[i], rng.choice(syms))] = 60000 + i
            gp[(rng.choice(syms), ids[-1 - i])] = 61000 + i
        gm = ora.Merges(gp)
        g = ctx.bpe_from_pairs(gp)
        gw = ora.run_buffer("bpe", data, eff, 2, gm)
        if not np.array_equal(g.tokenize_host(data, chunk_size=chunk or max(n, 1)), gw): bad.append("general map")
        g.close()
    cases += 1
    ok = not bad
    if not ok:
        fails += 1
        print("MISMATCH", index, bad, dict(variant=os.environ.get("BLT_SWEEP_VARIANT", "auto"), dense=os.environ["BLT_DENSE"], n=n, chunk=chunk, alpha=alpha,
                               style=style, density=density, rules=len(pairs), gap=gap), flush=True)
    s.close()
    ctx.close()
    if args.only >= 0:
        break
    index += 1
print(f"fuzz: {cases} cases, {fails} mismatches, seed {args.seed}")
sys.exit(1 if fails else 0)
