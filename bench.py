#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: BPE tokenize input GB/s (32k merges, 1 GiB synthetic text,
16 MiB chunks) on N B200s, next to the CPU path on the box's host cores.

A "step" is one pass of the hot path over one 1 GiB batch per GPU (configs[2] of BASELINE.json).
  value     device-resident: input already in HBM, output left in HBM (blt_process_resident)
  e2e       the same batch through the C-ABI call with pinned HOST buffers (blt_tokenize_host):
            per-chunk H2D, kernel and D2H inside the timed region
  roofline  algorithmic bytes (N_in + 2*T_out) / average kernel duration (CUDA events around every
            launch) against the measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline  the C++ oracle (restated reference, `kind: port`) on a bounded sample, all host threads
Multi-GPU: one process per GPU (torchrun), every rank tokenizes its own 1 GiB shard (weak scaling),
no data-path collective; only the timing is reduced (max over ranks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = 1 << 30
CHUNK = 16 << 20
N_MERGES = 32768
SEED = 0xB170003
METRIC = "BPE tokenize input GB/s at 1/2/4/8 B200 (32k merges, 1 GB) vs CPU ref"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_workload(n_bytes: int, rank: int, n_merges: int, out=None):
    """Rank r's shard: config-3 text with seed SEED + r; the merges table always comes from the first
    16 MiB of the rank-0 stream so that every rank uses the same table."""
    from blt_b200 import synth
    data = synth.text(n_bytes, SEED + rank, out=out)
    sample = data if rank == 0 else synth.text(min(n_bytes, synth.MERGE_SAMPLE_BYTES), SEED)
    left, right = synth.merges_from_sample(sample, n_merges)
    return data, left, right


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_port_rate(data, left, right, chunk, threads, target_s=12.0):
    """Times the oracle (restated reference CPU path) on a bounded prefix of the workload."""
    from oracle import oracle_ffi as ora
    import numpy as np
    pairs = {(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(left, right))}
    m = ora.Merges(pairs)
    probe = min(data.size, 4 * chunk)
    out = np.empty(2 * min(data.size, 64 * chunk) + 2, dtype=np.uint8)
    t0 = time.perf_counter()
    ora.run_buffer("bpe", data[:probe], chunk, threads, m, out_array=out)
    dt = time.perf_counter() - t0
    rate = probe / dt
    sample = int(min(data.size, max(probe, (rate * target_s) // chunk * chunk)))
    sample = min(sample, 64 * chunk)
    t0 = time.perf_counter()
    ora.run_buffer("bpe", data[:sample], chunk, threads, m, out_array=out)
    dt = time.perf_counter() - t0
    return sample / dt / 1e9, sample, dt


def run_reference(args):
    """The reference's own CPU implementation of the path.  Rust cannot be built in this image, so this
    is the C++ oracle port (labelled `kind: port`), multi-threaded like the reference's pipeline."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_ffi as ora
    import numpy as np
    threads = os.cpu_count() or 1
    n = args.bytes
    data, left, right = build_workload(min(n, 64 * CHUNK), 0, args.merges)
    pairs = {(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(left, right))}
    m = ora.Merges(pairs)
    out = np.empty(2 * data.size + 2, dtype=np.uint8)
    # size one step so that steps+warmup finish within a few minutes
    probe = min(data.size, 4 * CHUNK)
    t0 = time.perf_counter()
    ora.run_buffer("bpe", data[:probe], CHUNK, threads, m, out_array=out)
    rate = probe / (time.perf_counter() - t0)
    budget_s = 150.0 / max(1, args.steps + args.warmup)
    sample = int(max(CHUNK, min(data.size, (rate * budget_s) // CHUNK * CHUNK)))
    for _ in range(args.warmup):
        ora.run_buffer("bpe", data[:sample], CHUNK, threads, m, out_array=out)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ora.run_buffer("bpe", data[:sample], CHUNK, threads, m, out_array=out)
    dt = time.perf_counter() - t0
    gbs = sample * args.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 4), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8->u16", "data": "synthetic",
        "config": {"workload": f"BPE {args.merges} merges (u16 vocab), {n >> 20} MiB synthetic English-like text per GPU, "
                               f"16 MiB chunks, device-resident (BASELINE.json configs[2])",
                   "note": "reference arm: the same workload on the host cores (restated CPU path, in memory), "
                           "each step = a bounded prefix of it"},
        "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": f"first {sample >> 20} MiB of the workload per step, in memory, {threads} threads"},
        "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def pcie_probe(torch, h_in, h_out, d_in, d_out, n):
    """Pinned cudaMemcpyAsync ceilings on this box: H2D alone, D2H alone, both directions at once
    (what the e2e pipeline is bounded by: it moves n bytes in and the output back, concurrently)."""
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.wait_stream(torch.cuda.current_stream())
            s2.wait_stream(torch.cuda.current_stream())
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[:n].copy_(d_out[:n], non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1)
            torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return round(n / best / 1e6, 2)

    return {"h2d": run(True, False), "d2h": run(False, True), "both_directions_each": run(True, True)}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from blt_b200 import _native as nat

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: blt_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.bytes
    chunk = CHUNK
    n_chunks = (n + chunk - 1) // chunk

    # ---- workload: pinned host buffers (also used by the e2e leg), then device copies ----
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(2 * n + 2, dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter()
    data, left, right = build_workload(n, rank, args.merges, out=h_in.numpy())
    log(f"[rank {rank}] generated {n >> 20} MiB in {time.perf_counter() - t0:.1f}s")
    from blt_b200 import synth
    with tempfile.NamedTemporaryFile("w", suffix=".merges.txt", delete=False) as f:
        merges_path = f.name
    synth.write_merges_file(merges_path, left, right)
    ctx = nat.Context(local)
    strat = ctx.bpe_from_file(merges_path)      # merges.txt -> ids 256.., exactly as config_loader.rs
    os.unlink(merges_path)
    assert strat.num_merges == args.merges

    d_in = h_in.cuda(non_blocking=False)
    d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    d_ends = torch.zeros(n_chunks, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        strat.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), d_out.numel(), d_ends.data_ptr(),
                               stream, sync=False)

    for _ in range(max(args.warmup, 3)):
        step()
    out_bytes, sweeps = strat.resident_result(stream)
    t_out = out_bytes // 2

    # ---- parity spot check against the oracle (outside every timed region) ----
    parity = None
    if not args.no_check:
        from oracle import oracle_ffi as ora
        pairs = {(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(left, right))}
        om = ora.Merges(pairs)
        ends = d_ends.cpu().numpy()
        assert int(ends[-1]) == out_bytes
        for k in sorted({0, n_chunks - 1, n_chunks // 2, (7 * n_chunks) // 11}):
            lo = 0 if k == 0 else int(ends[k - 1])
            got = d_out[lo:int(ends[k])].cpu().numpy()
            want = np.frombuffer(ora.process_chunk("bpe", data[k * chunk:(k + 1) * chunk], om), dtype=np.uint8)
            if not np.array_equal(got, want):
                raise SystemExit(f"PARITY FAILURE in chunk {k}")
        parity = "oracle-checked chunks 0, mid, 7/11, last"

    # ---- timed region: device resident ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    time.sleep(0.25)          # let nvidia-smi attach; its first sample lands inside the timed region
    torch.cuda.synchronize()
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for a, b in evs:
        a.record()
        step()
        b.record()
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = t_begin.elapsed_time(t_end)
    kernel_ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)   # memset nodes + dense kernel + the three exact kernels (no-ops here), per step
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * n * args.steps / (total_ms * 1e-3) / 1e9

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e_steps = max(1, min(args.steps, 20))
        got = 0
        for _ in range(2):
            got = strat.tokenize_host_ptr(h_in.data_ptr(), n, chunk, h_out.data_ptr(), h_out.numel())
        assert got == out_bytes
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            got = strat.tokenize_host_ptr(h_in.data_ptr(), n, chunk, h_out.data_ptr(), h_out.numel())
        torch.cuda.synchronize()
        e_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        if not args.no_check:
            if not torch.equal(h_out[:got], d_out[:out_bytes].cpu()):
                raise SystemExit("PARITY FAILURE: e2e output differs from the device-resident output")
        pcie = pcie_probe(torch, h_in, h_out, d_in, d_out, n) if rank == 0 else None
        if world > 1:
            # every rank copies both ways at once: the host platform's aggregate ceiling for the e2e number
            # (on the pool's 8-GPU boxes ~72 GB/s per direction in total, far below 8 x one GPU's 50 GB/s)
            s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
                with torch.cuda.stream(s2):
                    h_out[:n].copy_(d_out[:n], non_blocking=True)
            s1.synchronize(); s2.synchronize()
            t = torch.tensor([(time.perf_counter() - t0) * 1e3], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                pcie["all_ranks_both_directions_each"] = round(world * 3 * n / (float(t.item()) * 1e-3) / 1e9, 2)
        e2e = {"value": round(world * n * e_steps / (e_ms * 1e-3) / 1e9, 3), "unit": "GB/s", "pcie_probe_GBps": pcie,
               "of_pcie_ceiling": None if not pcie else round(
                   world * n * e_steps / (e_ms * 1e-3) / 1e9 /
                   (pcie.get("all_ranks_both_directions_each") or pcie["both_directions_each"]), 3),
               "h2d_bytes_per_step": n, "d2h_bytes_per_step": int(out_bytes), "steps": e_steps,
               "ms_per_step": round(e_ms / e_steps, 3)}

    clocks = sampler.stop()   # sampled from just before the device-resident region to the end of the e2e region
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the sweep kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = n + 2 * t_out
    dense_held = (t_out == (n + 1) // 2)  # every token is a merged even pair <=> the dense pass's hypothesis held
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("bpe_sweep_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "kernel": ("bltk::dense_pairs_kernel (the only launch of a step: it held on every chunk, so the exact count/scan/emit "
                           "kernels it would launch from the device were not needed)") if dense_held else
                          "bltk::dense_pairs_kernel + the count/scan/emit kernels it launched from the device",
                "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms": round(kernel_ms, 4), "t_out_over_n_in": round(t_out / n, 4)}

    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        gbs, sample, dt = cpu_port_rate(data, left, right, chunk, threads)
        cpu_baseline = {"value": round(gbs, 4), "unit": "GB/s", "cores": threads, "kind": "port",
                        "sample": f"first {sample >> 20} MiB of the workload, in memory, {threads} threads, {dt:.1f}s"}

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8->u16", "data": "synthetic",
        "config": {"workload": f"BPE {args.merges} merges (u16 vocab), {n >> 20} MiB synthetic English-like text per GPU, "
                               f"16 MiB chunks, device-resident (BASELINE.json configs[2])",
                   "l2": "inputs larger than L2 (1 GiB in + out per step vs 126 MB L2), no flush needed",
                   "sweeps": sweeps, "variant": os.environ.get("BLT_SWEEP_VARIANT", "default"), "parity": parity},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": (1 if dense_held else 4) * args.steps,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=GIB)
    ap.add_argument("--merges", type=int, default=N_MERGES)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` typed by hand: start the N ranks the driver would start with torchrun
        import subprocess
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
