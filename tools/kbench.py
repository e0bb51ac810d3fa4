#!/usr/bin/env python
"""Kernel micro-benchmark: device-resident timings (CUDA events) of K1 (widen) and of every K2 tile
configuration on the BASELINE.json workloads.  Prints one JSON line per measurement.
    python tools/kbench.py [--bytes N] [--iters K] [--variants 0,1,2,3,4] [--configs 1,2,3,4,6,7,8,9]
(6 general map, 7 detokenizer, 8 pair histogram, 9 mixed corpus with 32 768 rules)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from blt_b200 import _native as nat, synth  # noqa: E402

if os.environ.get("BLT_ALT_LIB"):  # an alternative build of the library (A/B runs on one box)
    nat.LIB_PATH = os.environ["BLT_ALT_LIB"]

VARIANT_NAMES = ["exact r4", "exact r8", "exact walk", "fused 15x4 d2", "fused 23x2 d2"]


def time_resident(strat, d_in, n, chunk, d_out, iters):
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        strat.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), d_out.numel(), 0, stream, sync=False)
    out_len, _ = strat.resident_result(stream)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    torch.cuda.synchronize()
    for a, b in evs:
        a.record()
        strat.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), d_out.numel(), 0, stream, sync=False)
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    return out_len, ms[len(ms) // 2], ms[0]


def general_map(peak):
    """K3: a map merges.txt cannot express (second-level rules on produced ids) -> hash front end, several
    sweeps per chunk with a host round trip after each (SURVEY.md 8d, config 4's 64 MiB slice)."""
    import time
    n, chunk = 64 << 20, 16 << 20
    data = synth.text(n, synth.SEED_CONFIG[4])
    l, r = synth.merges_from_sample(data, 256)
    pairs = {(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))}
    for i in range(64):
        pairs[(256 + i, 32)] = 1000 + i        # (merged id, space) -> second-level id
        pairs[(1000 + i, 256 + i)] = 2000 + i  # third level
    ctx = nat.Context(0)
    strat = ctx.bpe_from_pairs(pairs)
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    times = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out_len = strat.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), d_out.numel(), 0, stream, sync=True)
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    _, sweeps = strat.resident_result(stream)
    med = sorted(times)[len(times) // 2]
    alg = n + out_len
    print(json.dumps({"config": "general map (K3, hash front end)", "n": n, "rules": len(pairs), "sweeps": sweeps, "out_bytes": out_len,
                      "ms_median_wall": round(med, 3), "input_GBps": round(n / med / 1e6, 1),
                      "algorithmic_GBps": round(alg / med / 1e6, 1), "frac_of_measured_hbm": round(alg / med / 1e6 / peak, 4)}), flush=True)
    strat.close()
    ctx.close()


def detok(peak, n, chunk, iters):
    """Detokenizer on the token streams of configs 3 (every token a merged id), 2 (mixed widths) and of
    the mixed corpus with 32 768 rules (9)."""
    stream = torch.cuda.current_stream().cuda_stream
    for cfg, rules in ((3, 32768), (2, 256), (9, 32768)):
        data = synth.mixed(n, synth.SEED_MIXED) if cfg == 9 else synth.text(n, synth.SEED_CONFIG[cfg])
        l, r = synth.merges_from_sample(data, rules)
        ctx = nat.Context(0)
        strat = ctx.bpe_from_pairs({(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))})
        d_in = torch.from_numpy(data).cuda()
        d_tok = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
        nt = strat.process_resident(d_in.data_ptr(), n, chunk, d_tok.data_ptr(), 2 * n, 0, stream, sync=True)
        d_back = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            got = strat.detokenize_resident(d_tok.data_ptr(), nt, d_back.data_ptr(), n, stream, sync=True)
        assert got == n and torch.equal(d_back[:n], d_in)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in evs:
            a.record()
            strat.detokenize_resident(d_tok.data_ptr(), nt, d_back.data_ptr(), n, stream, sync=False)
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        med = ms[len(ms) // 2]
        alg = nt + n
        print(json.dumps({"config": f"detokenize config {cfg} tokens", "form": {"0": "count/scan/emit", "1": "fused"}.get(os.environ.get("BLT_DETOK_VARIANT", "0"), "?"), "token_bytes": nt, "out_bytes": n, "ms_median": round(med, 4),
                          "ms_best": round(ms[0], 4), "output_GBps": round(n / med / 1e6, 1), "algorithmic_GBps": round(alg / med / 1e6, 1),
                          "frac_of_measured_hbm": round(alg / med / 1e6 / peak, 4)}), flush=True)
        strat.close()
        ctx.close()
        del d_in, d_tok, d_back


def pair_hist(peak, n, iters):
    """Adjacent-byte-pair histogram on text and on uniform random bytes (every counter hot)."""
    stream = torch.cuda.current_stream().cuda_stream
    ctx = nat.Context(0)
    for name, data in (("text", synth.text(n, synth.SEED_CONFIG[3])), ("random bytes", synth.random_bytes(n, synth.SEED_CONFIG[1]))):
        d_in = torch.from_numpy(data).cuda()
        d_counts = torch.zeros(65536, dtype=torch.int64, device="cuda")
        for _ in range(2):
            ctx.count_pairs_resident(d_in.data_ptr(), n, d_counts.data_ptr(), stream)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in evs:
            a.record()
            ctx.count_pairs_resident(d_in.data_ptr(), n, d_counts.data_ptr(), stream)
            b.record()
        torch.cuda.synchronize()
        assert int(d_counts.sum().item()) == n - 1
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        med = ms[len(ms) // 2]
        print(json.dumps({"config": f"pair histogram, {name}", "n": n, "ms_median": round(med, 4), "ms_best": round(ms[0], 4),
                          "input_GBps": round(n / med / 1e6, 1), "frac_of_measured_hbm": round(n / med / 1e6 / peak, 4)}), flush=True)
        del d_in
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--variants", default="auto")
    ap.add_argument("--configs", default="1,2,3,4")
    args = ap.parse_args()
    n, chunk = args.bytes, 16 << 20
    peak = 6555.5
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    torch.cuda.set_device(0)
    d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    for cfg in [int(c) for c in args.configs.split(",")]:
        if cfg == 6:
            general_map(peak)
            continue
        if cfg == 7:
            detok(peak, n, chunk, args.iters)
            continue
        if cfg == 8:
            pair_hist(peak, n, args.iters)
            continue
        if cfg == 1:
            data = synth.random_bytes(n, synth.SEED_CONFIG[1])
        elif cfg == 4:
            data = synth.adversarial(n, synth.SEED_CONFIG[4])
        elif cfg == 9:   # mixed corpus: every byte pair occurs, the 32 768-rule table is a strict subset (the exact sweep's real workload)
            data = synth.mixed(n, synth.SEED_MIXED)
        else:
            data = synth.text(n, synth.SEED_CONFIG[cfg])
        d_in = torch.from_numpy(data).cuda()
        variants = [None] if cfg == 1 else [(-1 if v == "auto" else int(v)) for v in args.variants.split(",")]
        for v in variants:
            if v is not None and v >= 0:
                os.environ["BLT_SWEEP_VARIANT"] = str(v)
            else:
                os.environ.pop("BLT_SWEEP_VARIANT", None)   # the library chooses (fused, or three launches on all-merging input)
            ctx = nat.Context(0)
            if cfg == 1:
                strat, name = ctx.basic(), "widen"
            elif cfg == 4:
                strat = ctx.bpe_from_pairs({p_: 256 + i for i, p_ in enumerate(synth.adversarial_pairs())})
                name = VARIANT_NAMES[v] if v >= 0 else "auto"
            else:
                l, r = synth.merges_from_sample(data, 256 if cfg == 2 else 32768)
                strat = ctx.bpe_from_pairs({(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))})
                name = VARIANT_NAMES[v] if v >= 0 else "auto"
            if cfg != 1:
                name = ("dense+" if os.environ.get("BLT_DENSE", "1") != "0" else "") + name
            out_len, med, best = time_resident(strat, d_in, n, chunk, d_out, args.iters)
            alg = n + out_len
            print(json.dumps({"config": cfg, "kernel": name, "n": n, "out_bytes": out_len, "ratio_tokens_per_byte": round(out_len / 2 / n, 4),
                              "ms_median": round(med, 4), "ms_best": round(best, 4), "input_GBps": round(n / med / 1e6, 1),
                              "algorithmic_GBps": round(alg / med / 1e6, 1), "frac_of_measured_hbm": round(alg / med / 1e6 / peak, 4)}),
                  flush=True)
            strat.close()
            ctx.close()
        del d_in


if __name__ == "__main__":
    main()
