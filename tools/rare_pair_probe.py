#!/usr/bin/env python
"""What ONE even pair that is not a rule costs per GiB ("speculation granularity"): configs[2] text with a
single foreign byte in the middle, (a) device-resident in one 1 GiB launch, (b) end to end through blt_tokenize_host
(units of <= 64 MiB, pinned buffers).  Prints clean vs poisoned times; the output is checked against the oracle."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from blt_b200 import _native as nat, synth
from oracle import oracle_ffi as ora

n, chunk = 1 << 30, 16 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(2 * n + 2, dtype=torch.uint8).pin_memory()
data = synth.text(n, synth.SEED_CONFIG[3], out=h_in.numpy())
l, r = synth.merges_from_sample(data, 32768)
pairs = {(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))}
ctx = nat.Context(0)
stream = torch.cuda.current_stream().cuda_stream
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
res = {}
for label in ("clean", "one foreign byte per GiB"):
    if label != "clean":
        data[n // 2 + 12346] = 255        # even offset in its chunk: (0xFF, y) is not a rule (the padding of the table covers low keys only)
    d_in.copy_(h_in)
    s = ctx.bpe_from_pairs(pairs)          # a fresh predictor
    want = ora.run_buffer("bpe", data, chunk, os.cpu_count() or 4, ora.Merges(pairs))
    ms = []
    for i in range(40):                    # the predictor's steady state: after a failure most calls skip the dense attempt
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out_len = s.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), 2 * n, 0, stream)
        torch.cuda.synchronize(); ms.append((time.perf_counter() - t0) * 1e3)
    assert out_len == want.size and np.array_equal(d_out[:out_len].cpu().numpy(), want)
    assert (out_len == n) == (label == "clean"), "the foreign byte must break the dense hypothesis"
    e = []
    for i in range(6):
        t0 = time.perf_counter()
        got = s.tokenize_host_ptr(h_in.data_ptr(), n, chunk, h_out.data_ptr(), h_out.numel())
        e.append((time.perf_counter() - t0) * 1e3)
    assert got == want.size and np.array_equal(h_out[:got].numpy(), want)
    res[label] = {"resident_ms_first_call": round(ms[0], 3), "resident_ms_median_of_40": round(sorted(ms)[20], 3),
                  "e2e_ms_best_of_6": round(min(e), 2), "e2e_GBps": round(n / min(e) / 1e6, 2)}
    s.close()
print(json.dumps(res))
