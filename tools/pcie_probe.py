#!/usr/bin/env python
"""Pinned-memory PCIe probe: H2D alone, D2H alone, both at once (the e2e pipeline's ceiling)."""
import json, torch
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5, chunk=None):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        step = chunk or n
        for off in range(0, n, step):
            if h2d:
                with torch.cuda.stream(s1): d_a[off:off+step].copy_(h_in[off:off+step], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out[off:off+step].copy_(d_b[off:off+step], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return round(n / best / 1e6, 2)
print(json.dumps({"h2d_GBps": run(True, False), "d2h_GBps": run(False, True), "both_each_GBps": run(True, True),
                  "both_each_16MiB_copies_GBps": run(True, True, chunk=16 << 20)}))
