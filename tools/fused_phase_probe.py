#!/usr/bin/env python
"""Where the workers of the fused sweep spend their cycles (needs a library built with -DBLT_FUSED_PROF, path in
BLT_PROF_LIB): waiting for the slice copy / counting / waiting for the chain warp / emitting, per worker and SM."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from blt_b200 import _native as nat, synth  # noqa: E402

nat.LIB_PATH = os.environ["BLT_PROF_LIB"]
os.environ["BLT_SWEEP_VARIANT"] = os.environ.get("BLT_SWEEP_VARIANT", "3")
os.environ["BLT_DENSE"] = "0"
n, chunk = 1 << 30, 16 << 20
for cfg in (2, 3):
    data = synth.text(n, synth.SEED_CONFIG[cfg])
    l, r = synth.merges_from_sample(data, 256 if cfg == 2 else 32768)
    ctx = nat.Context(0)
    s = ctx.bpe_from_pairs({(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(l, r))})
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        s.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), 2 * n, 0, st, sync=True)
    buf = np.zeros(256 * 32 * 8, dtype=np.uint64)
    assert nat.lib().blt_debug_fused_profile(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size)) == 0
    ch = buf.reshape(256, 32, 8).astype(np.float64)[:148, 31]
    c2 = buf.reshape(256, 32, 8).astype(np.float64)[:148, 30]
    print(json.dumps({"config": cfg, "chain_warp_per_tile": {"first_to_last_worker_cycles": round(float((ch[:, 0] / ch[:, 4]).mean())),
                      "publish_to_pickup_cycles": round(float((ch[:, 1] / ch[:, 4]).mean())),
                      "publish_to_resolved_cycles": round(float((ch[:, 2] / ch[:, 4]).mean())),
                      "polls": round(float((ch[:, 3] / ch[:, 4]).mean()), 2),
                      "cycles_per_poll": round(float((c2[:, 0] / ch[:, 3]).mean())),
                      "first_poll_unpublished_tiles": round(float((c2[:, 1] / c2[:, 4]).mean()), 2),
                      "first_poll_distance_to_prefix": round(float((c2[:, 2] / c2[:, 4]).mean()), 1),
                      "first_poll_cycles_after_publish": round(float((c2[:, 3] / c2[:, 4]).mean()))}}), flush=True)
    W = 15 if os.environ["BLT_SWEEP_VARIANT"] == "3" else 23
    p = buf.reshape(256, 32, 8).astype(np.float64)[:148, :W]
    tot = p[:, :, 0:4].sum(axis=2)
    frac = p[:, :, 0:4] / tot[:, :, None]
    print(json.dumps({"config": cfg, "mean_frac_copywait_count_chainwait_emit": [round(float(x), 4) for x in frac.mean(axis=(0, 1))],
                      "per_worker_chainwait": [round(float(x), 3) for x in frac[:, :, 2].mean(axis=0)],
                      "per_worker_count_cycles_per_tile": [round(float(x)) for x in (p[:, :, 1] / p[:, :, 4]).mean(axis=0)],
                      "per_worker_emit_cycles_per_tile": [round(float(x)) for x in (p[:, :, 3] / p[:, :, 4]).mean(axis=0)],
                      "cta_chainwait_min_med_max": [round(float(x), 3) for x in np.percentile(frac[:, :, 2].mean(axis=1), [0, 50, 100])],
                      "cta_count_cycles_min_med_max": [round(float(x)) for x in np.percentile((p[:, :, 1] / p[:, :, 4]).mean(axis=1), [0, 50, 100])],
                      "total_cycles_per_tile": round(float((tot / p[:, :, 4]).mean()))}), flush=True)
    s.close(); ctx.close()
