#!/usr/bin/env python
"""Aggregate host<->device bandwidth with G GPUs copying at once (pinned memory, both directions):
what bounds bench.py's e2e number at N > 1.  One process per GPU, barrier-synchronised, 3 s each."""
import json, os, sys, time
import multiprocessing as mp


def worker(g, G, barrier, q, seconds, numa):
    os.environ["CUDA_VISIBLE_DEVICES"] = str(g)
    import torch
    if numa:
        try:
            p = torch.cuda.get_device_properties(0)
            path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/numa_node"
            node = int(open(path).read())
            if node >= 0:
                cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
                ids = []
                for part in cpus.split(","):
                    a, _, b = part.partition("-")
                    ids += list(range(int(a), int(b or a) + 1))
                os.sched_setaffinity(0, set(ids) & os.sched_getaffinity(0) or os.sched_getaffinity(0))
        except Exception as e:  # noqa: BLE001
            print("numa:", e, file=sys.stderr)
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for name, h2d, d2h in (("h2d", True, False), ("d2h", False, True), ("both_each", True, True)):
        torch.cuda.synchronize(); barrier.wait()
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < seconds:
            if h2d:
                with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
            s1.synchronize(); s2.synchronize(); reps += 1
        res[name] = reps * n / (time.perf_counter() - t0) / 1e9
    q.put((g, res))


if __name__ == "__main__":
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    numa = "--numa" in sys.argv
    mp.set_start_method("spawn")
    barrier, q = mp.Barrier(G), mp.Queue()
    ps = [mp.Process(target=worker, args=(g, G, barrier, q, 3.0, numa)) for g in range(G)]
    [p.start() for p in ps]
    out = dict(q.get() for _ in ps)
    [p.join() for p in ps]
    agg = {k: round(sum(out[g][k] for g in out), 1) for k in ("h2d", "d2h", "both_each")}
    print(json.dumps({"gpus": G, "numa_affinity": numa, "aggregate_GBps": agg,
                      "per_gpu_both_each": [round(out[g]["both_each"], 1) for g in sorted(out)]}))
