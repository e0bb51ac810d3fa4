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
  exact     a second, non-degenerate workload (mixed corpus: > 200 distinct bytes, every pair occurs, the 32 768-rule
            table is a strict subset of the observed pairs): the dense speculation never holds there, so this is the
            exact single-pass sweep (greedy leftmost run parity + look-back compaction): roofline_exact,
            dense_hit_rate, t_out_over_n_in
  e2e_chunk_api  blt_process_chunk (the reference's TokenizationStrategy::process_chunk) on pageable 16 MiB slices
            from 1 and 8 calling threads, as pipeline.rs:141-150 would call it
  file_to_file  blt_run_tokenizer (pipeline.rs:56-131) tmpfs file -> tmpfs file on N GPUs: configs[2] and, when the
            box has the memory, configs[4] (8 GiB, 60 000 rules); input GB/s, wall, fraction of the PCIe roofline
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


def load_synth():
    """blt_b200/synth.py as a stand-alone module (ctypes over libblt_synth.so): importing the package would map
    libblt_cuda.so, which the reference arm must not touch."""
    import importlib.util
    if "blt_bench_synth" in sys.modules:
        return sys.modules["blt_bench_synth"]
    spec = importlib.util.spec_from_file_location("blt_bench_synth", os.path.join(ROOT, "blt_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["blt_bench_synth"] = mod
    spec.loader.exec_module(mod)
    return mod


def build_workload(n_bytes: int, rank: int, n_merges: int, out=None, kind: str = "text"):
    """Rank r's shard: config-3 text (or the mixed corpus) with seed + r; the merges table always comes from the
    first 16 MiB of the rank-0 stream so that every rank uses the same table."""
    synth = load_synth()
    gen, seed = (synth.text, SEED) if kind == "text" else (synth.mixed, synth.SEED_MIXED)
    data = gen(n_bytes, seed + rank, out=out)
    sample = data if rank == 0 else gen(min(n_bytes, synth.MERGE_SAMPLE_BYTES), seed)
    left, right = synth.merges_from_sample(sample, n_merges)
    return data, left, right


def workload_config(args):
    """The `config` object, identical in both arms."""
    return {"workload": f"BPE {args.merges} merges (u16 vocab), {args.bytes >> 20} MiB synthetic English-like text per GPU, "
                        f"16 MiB chunks (BASELINE.json configs[2])"}


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
        "config": workload_config(args),
        "notes": {"arm": "the same workload on the host cores (restated CPU path, in memory), each step = a bounded prefix of it",
                  "maps_libblt_cuda": any("libblt_cuda" in ln for ln in open("/proc/self/maps"))},
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


def check_against_oracle(np, data, left, right, chunk, d_out, d_ends, out_bytes):
    """Whole-output comparison with the oracle (multi-threaded C++ restatement): every byte and every chunk end."""
    from oracle import oracle_ffi as ora
    pairs = {(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(left, right))}
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    want = ora.run_buffer("bpe", data, chunk, threads, ora.Merges(pairs))
    dt = time.perf_counter() - t0
    got = d_out[:out_bytes].cpu().numpy()
    if want.size != out_bytes or not np.array_equal(got, want):
        raise SystemExit("PARITY FAILURE: device-resident output differs from the oracle")
    ends = d_ends.cpu().numpy()
    if int(ends[-1]) != out_bytes or np.any(np.diff(ends) <= 0):
        raise SystemExit("PARITY FAILURE: chunk ends")
    return f"whole output ({out_bytes} bytes) and {ends.size} chunk ends equal to the oracle's ({dt:.1f}s on {threads} threads)"


def time_resident(torch, strat, d_in, n, chunk, d_out, d_ends, stream, steps, warmup):
    """`steps` launches of blt_process_resident with CUDA events around each; returns (mean ms per launch, out_bytes)."""
    for _ in range(warmup):
        strat.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), d_out.numel(), d_ends.data_ptr(), stream, sync=False)
    out_bytes, _ = strat.resident_result(stream)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    for a, b in evs:
        a.record()
        strat.process_resident(d_in.data_ptr(), n, chunk, d_out.data_ptr(), d_out.numel(), d_ends.data_ptr(), stream, sync=False)
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / len(evs), out_bytes


def chunk_api_leg(np, strat, data, chunk, want, threads_list=(1, 8)):
    """blt_process_chunk from T host threads on pageable slices of one buffer (the reference's mmap) into fresh
    pageable outputs (its Vec<u8>), results consumed in chunk order: what `impl TokenizationStrategy` costs."""
    n = data.size
    n_chunks = (n + chunk - 1) // chunk
    pageable = np.array(data, copy=True)          # NOT the pinned buffer: the reference hands in slices of an mmap
    res = {}
    for T in threads_list:
        best, outs = None, None
        for _ in range(2):
            outs = [None] * n_chunks
            nxt = [0]
            lock = threading.Lock()

            def worker():
                while True:
                    with lock:
                        k = nxt[0]
                        nxt[0] += 1
                    if k >= n_chunks:
                        return
                    buf = np.empty(2 * chunk, dtype=np.uint8)
                    outs[k] = strat.process_chunk(pageable[k * chunk:(k + 1) * chunk], out=buf)

            t0 = time.perf_counter()
            ths = [threading.Thread(target=worker) for _ in range(T)]
            [t.start() for t in ths]
            [t.join() for t in ths]
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        if want is not None and not np.array_equal(np.concatenate(outs), want):
            raise SystemExit("PARITY FAILURE: blt_process_chunk output differs")
        res[f"callers_{T}"] = round(n / best / 1e9, 2)
    return {"unit": "GB/s", **res, "buffers": "pageable input slices, fresh pageable output per chunk", "chunk_bytes": chunk,
            "h2d_bytes_per_step": int(n), "d2h_bytes_per_step": int(want.size) if want is not None else None}


def fresh_page_rate(d, n=1 << 30):
    """How fast this box produces fresh, mapped pages of a new file in `d` (fallocate + MADV_POPULATE_WRITE from one
    thread, what the file pipeline's background thread does): the ceiling of any mapped fresh output file here."""
    import ctypes, mmap
    libc = ctypes.CDLL(None, use_errno=True)
    path = os.path.join(d, f"blt_bench_pages_{os.getpid()}")
    fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        os.ftruncate(fd, n)
        m = mmap.mmap(fd, n)
        addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
        t0 = time.perf_counter()
        os.posix_fallocate(fd, 0, n)
        libc.madvise(ctypes.c_void_p(addr), ctypes.c_size_t(n), 23)   # MADV_POPULATE_WRITE
        dt = time.perf_counter() - t0
        del addr
        try:
            m.close()
        except BufferError:
            pass
        return round(n / dt / 1e9, 2)
    finally:
        os.close(fd)
        os.unlink(path)


def file_leg(np, nat, synth, name, data, left, right, n_gpus, pcie, want=None, reps=3, page_gbps=None):
    """blt_run_tokenizer tmpfs -> tmpfs on n_gpus GPUs of this process; wall = the whole call (config, mmap, contexts
    on first use, pipeline, trim).  frac_of_pcie_roofline = max(N_in / BW_h2d, out / BW_d2h) / pipeline time."""
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    inp, outp, mp = (os.path.join(d, f"blt_bench_{os.getpid()}_{name}.{e}") for e in ("in", "out", "merges.txt"))
    try:
        data.tofile(inp)
        synth.write_merges_file(mp, left, right)
        walls = []
        for _ in range(reps):
            t0 = time.perf_counter()
            nat.run_tokenizer(inp, outp, merges_file=mp, chunk_size="16MB", num_gpus=n_gpus)
            walls.append(time.perf_counter() - t0)
        out_bytes = os.path.getsize(outp)
        if want is not None:
            got = np.memmap(outp, dtype=np.uint8, mode="r")
            if got.size != want.size or not np.array_equal(got, want):
                raise SystemExit(f"PARITY FAILURE: file-to-file output of {name} differs")
            del got
        best = min(walls)
        bw_in = pcie["h2d"] * (n_gpus if n_gpus > 1 and "all_ranks_both_directions_each" not in pcie else 1)
        bw_out = pcie["d2h"] * (n_gpus if n_gpus > 1 and "all_ranks_both_directions_each" not in pcie else 1)
        if "all_ranks_both_directions_each" in pcie:
            bw_in = bw_out = pcie["all_ranks_both_directions_each"]
        t_roof = max(data.size / (bw_in * 1e9), out_bytes / (bw_out * 1e9))
        return {"workload": name, "gpus": n_gpus, "bytes_in": int(data.size), "bytes_out": int(out_bytes),
                "wall_s_first_call": round(walls[0], 3), "wall_s_best": round(best, 3), "input_GBps": round(data.size / best / 1e9, 2),
                "pcie_roofline_s": round(t_roof, 4), "frac_of_pcie_roofline": round(t_roof / best, 3), "filesystem": d,
                "fresh_output_page_GBps": page_gbps,
                "frac_of_output_page_ceiling": None if not page_gbps else round(out_bytes / (page_gbps * 1e9) / best, 3)}
    finally:
        for f in (inp, outp, mp):
            if os.path.exists(f):
                os.unlink(f)


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
    synth = load_synth()
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

    # ---- parity against the oracle: the WHOLE output and every chunk end (outside every timed region) ----
    parity = None
    if not args.no_check:
        parity = check_against_oracle(np, data, left, right, chunk, d_out, d_ends, out_bytes)

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
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return   # rank 0 alone runs the informational legs below (chunk API, exact workload, file to file on all N GPUs)

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

    # ---- the drop-in call itself: blt_process_chunk on pageable slices (rank 0, its own GPU) ----
    e2e_chunk_api = None
    if not args.no_e2e:
        want = h_out[:out_bytes].numpy() if e2e is not None else None
        e2e_chunk_api = chunk_api_leg(np, strat, data, chunk, want)

    # ---- non-degenerate workload: the exact sweep ----
    exact = None
    if not args.no_exact:
        t0 = time.perf_counter()
        mdata, mleft, mright = build_workload(n, 0, args.merges, out=h_in.numpy(), kind="mixed")   # reuses the pinned input buffer
        log(f"[rank 0] generated {n >> 20} MiB of mixed corpus in {time.perf_counter() - t0:.1f}s")
        mstrat = ctx.bpe_from_pairs({(int(a), int(b)): 256 + i for i, (a, b) in enumerate(zip(mleft, mright))})
        d_in.copy_(h_in)
        x_steps = max(3, min(args.steps, 20))
        x_ms, x_out = time_resident(torch, mstrat, d_in, n, chunk, d_out, d_ends, stream, x_steps, 3)
        x_tok = x_out // 2
        ends = np.concatenate([[0], d_ends.cpu().numpy()])
        lens_in = np.minimum(chunk, n - chunk * np.arange(n_chunks))
        dense_chunks = int(np.sum(np.diff(ends) == 2 * ((lens_in + 1) // 2)))
        x_parity = None
        if not args.no_check:
            x_parity = check_against_oracle(np, mdata, mleft, mright, chunk, d_out, d_ends, x_out)
        x_alg = n + 2 * x_tok
        x_ach = x_alg / (x_ms * 1e-3) / 1e9
        xtraffic = None
        if os.path.exists(tpath):
            try:
                xtraffic = json.load(open(tpath)).get("exact_sweep_dram_bytes_per_launch")
            except Exception:
                xtraffic = None
        exact = {"workload": f"BPE {args.merges} merges on {n >> 20} MiB of a mixed corpus (text, code, UTF-8 script, base64, binary; "
                             "every byte pair occurs, the table is a strict subset), 16 MiB chunks, device-resident",
                 "value": round(n / (x_ms * 1e-3) / 1e9, 2), "unit": "GB/s", "ms_per_step": round(x_ms, 4), "steps": x_steps,
                 "t_out_over_n_in": round(x_tok / n, 4), "dense_hit_rate": round(dense_chunks / n_chunks, 4),
                 "roofline_exact": {"bound": "hbm", "achieved": round(x_ach, 1), "peak": peak, "unit": "GB/s", "frac": round(x_ach / peak, 4),
                                    "traffic": xtraffic, "algorithmic_bytes_per_launch": int(x_alg),
                                    "kernel": "bltk::fused_sweep_kernel (count + decoupled look-back + emit in one pass) behind the dense "
                                              "pass's predictor (which stops attempting after the first failures)"},
                 "parity": x_parity}
        mstrat.close()

    # ---- file to file on all N GPUs from this process (the other ranks have left) ----
    file_to_file = None
    if not args.no_file and e2e is not None and e2e.get("pcie_probe_GBps"):
        pcie = e2e["pcie_probe_GBps"]
        del d_in, d_out, d_ends
        torch.cuda.empty_cache()
        file_to_file = []
        try:
            page_gbps = fresh_page_rate("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
        except Exception as e:  # informational only
            log(f"fresh page probe failed: {e}")
            page_gbps = None
        tdata, tleft, tright = build_workload(n, 0, args.merges, out=h_in.numpy())      # the headline workload again
        file_to_file.append(file_leg(np, nat, synth, "configs2_text_32768_merges", tdata, tleft, tright, world, pcie,
                                     want=None if args.no_check else h_out[:out_bytes].numpy(), page_gbps=page_gbps))
        try:
            import psutil
            free = psutil.virtual_memory().available
            shm_free = os.statvfs("/dev/shm").f_bavail * os.statvfs("/dev/shm").f_frsize if os.path.isdir("/dev/shm") else 0
        except Exception:
            free = shm_free = 0
        n5 = 8 * GIB
        if min(free, shm_free) > 5 * n5 and not args.no_config5:
            t0 = time.perf_counter()
            d5 = synth.text(n5, synth.SEED_CONFIG[5])
            l5, r5 = synth.merges_from_sample(d5, 60000)
            log(f"[rank 0] generated config 5 (8 GiB) in {time.perf_counter() - t0:.1f}s")
            file_to_file.append(file_leg(np, nat, synth, "configs4_8GiB_60000_merges", d5, l5, r5, world, pcie, reps=2, page_gbps=page_gbps))
            del d5
        else:
            file_to_file.append({"workload": "configs4_8GiB_60000_merges", "skipped": f"needs {5 * n5 >> 30} GiB of free RAM and tmpfs "
                                 f"(have {free >> 30} / {shm_free >> 30} GiB); tests/test_gpu_parity.py::test_config5_file_to_file covers it"})

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8->u16", "data": "synthetic",
        "config": workload_config(args),
        "notes": {"value": "device-resident: input and output stay in HBM (blt_process_resident)",
                  "l2": "inputs larger than L2 (1 GiB in + out per step vs 126 MB L2), no flush needed",
                  "sweeps": sweeps, "variant": os.environ.get("BLT_SWEEP_VARIANT", "default"), "parity": parity},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "exact": exact, "e2e_chunk_api": e2e_chunk_api,
        "file_to_file": file_to_file, "gpu_launches": (1 if dense_held else 3) * args.steps,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--no-exact", action="store_true")
    ap.add_argument("--no-file", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
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
