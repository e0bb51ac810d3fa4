#!/usr/bin/env python
"""Debug: builds a -DBLT_TRACE copy of the library, runs the config-3 sweep and prints per-phase
durations (globaltimer ns) of warp 0 / group 0 of every CTA.  Not part of the product."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
csrc = os.path.join(ROOT, "blt_b200", "csrc")
lib = os.path.join(ROOT, "blt_b200", "lib", "libblt_trace.so")
if not os.path.exists(lib):
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-DBLT_TRACE",
                    "-Xcompiler", "-fPIC,-fvisibility=hidden,-pthread", "-o", lib] +
                   [os.path.join(csrc, f) for f in ("kernels.cu", "cabi.cu", "pipeline.cu", "host_config.cpp")], check=True)
from blt_b200 import _native as nat, synth
nat.LIB_PATH = lib
nat._lib = None
variant = sys.argv[1] if len(sys.argv) > 1 else "2"
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
os.environ["BLT_SWEEP_VARIANT"] = variant
n = 1 << 30
data = synth.text(n, synth.SEED_CONFIG[cfg])
l, r = synth.merges_from_sample(data, 32768 if cfg == 3 else 256)
ctx = nat.Context(0)
s = ctx.bpe_from_pairs({(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))})
d_in = torch.from_numpy(data).cuda(); d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
ITERS = 256
trace = torch.zeros(148 * ITERS * 8, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    s.process_resident(d_in.data_ptr(), n, 16 << 20, d_out.data_ptr(), 2 * n, 0, st, sync=True)
L = nat.lib(); L.blt_debug_set_trace.argtypes = [C.c_void_p, C.c_uint]
assert L.blt_debug_set_trace(trace.data_ptr(), ITERS) == 0
s.process_resident(d_in.data_ptr(), n, 16 << 20, d_out.data_ptr(), 2 * n, 0, st, sync=True)
t = trace.cpu().numpy().reshape(148, ITERS, 8).astype(np.int64)
valid = (t[:, :, 0] > 0) & (t[:, :, 6] > 0) & (t[:, :, 4] > 0)
names = ["A1 (own warp)", "wait barrier A", "publish AGG", "look-back", "to barrier B passed", "emit+prefetch"]
print("variant", variant, "config", cfg, "valid iterations", int(valid.sum()))
for i, nm in enumerate(names):
    d = (t[:, :, i + 1] - t[:, :, i])[valid]
    print(f"{nm:22s} mean {d.mean():8.0f} ns   p50 {np.median(d):8.0f}   p90 {np.percentile(d, 90):8.0f}   max {d.max():8.0f}")
per = (t[:, 1:, 0] - t[:, :-1, 0])[valid[:, 1:] & valid[:, :-1]]
print(f"{'iteration period':22s} mean {per.mean():8.0f} ns   p50 {np.median(per):8.0f}")
# skew between CTAs at the same iteration
it0 = t[:, 5, 0]; print("start skew across CTAs at iter 5: ", int(it0[it0 > 0].max() - it0[it0 > 0].min()), "ns")

lb = (t[:, :, 4] - t[:, :, 3]).astype(np.float64); lb[~valid] = np.nan
print("look-back mean by CTA index (deciles):", [int(np.nanmean(lb[i:i + 15])) for i in range(0, 148, 15)])
for it in (5, 50, 150):
    st0 = t[:, it, 0].astype(np.float64); st0[st0 == 0] = np.nan
    rel = st0 - np.nanmin(st0)
    print(f"iter {it}: start offset by CTA (ns), every 15th CTA:", [int(rel[i]) if not np.isnan(rel[i]) else -1 for i in range(0, 148, 15)])
    agg = t[:, it, 3].astype(np.float64); agg[agg == 0] = np.nan
    pre = t[:, it, 4].astype(np.float64); pre[pre == 0] = np.nan
    print(f"   AGG publish offset:", [int(agg[i] - np.nanmin(st0)) if not np.isnan(agg[i]) else -1 for i in range(0, 148, 15)])
    print(f"   look-back done offset:", [int(pre[i] - np.nanmin(st0)) if not np.isnan(pre[i]) else -1 for i in range(0, 148, 15)])
dbg = t[:, :, 7][valid]
print("polls per look-back: mean", (dbg & 0xFFFF).mean(), " not-ready retries: mean", ((dbg >> 16) & 0xFFFF).mean(), " general-fold fraction", ((dbg >> 32) & 1).mean())
