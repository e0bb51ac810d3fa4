#!/usr/bin/env python
"""Randomised file-to-file fuzzing of the `blt` CLI (single and multi GPU, files and pipes) against the oracle.
python tools/fuzz_cli.py --seconds 100 --max-gpus 2"""
import argparse, os, random, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from blt_b200 import synth
from oracle import oracle_ffi as ora

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=100)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--max-gpus", type=int, default=1)
args = ap.parse_args()
BLT = os.path.join(ROOT, "blt_b200", "lib", "blt")
TYPES = {"text": 0xFF01, "audio": 0xFF02, "bin": 0xFF03, "video": 0xFF04}
t_end = time.time() + args.seconds
cases = fails = 0
with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
    idx = 0
    while time.time() < t_end:
        rng = random.Random(args.seed * 7919 + idx)
        idx += 1
        n = rng.choice([0, 1, rng.randint(2, 5000), rng.randint(5000, 2 << 20), rng.randint(2 << 20, 40 << 20)])
        data = synth.text(n, rng.randrange(1 << 30)) if n else np.zeros(0, np.uint8)
        mode = rng.choice(["basic", "bpe", "bpe", "passthrough"])
        cs_arg, eff = rng.choice([(None, 16 << 20), ("256KB", 256 << 10), ("1MB", 1 << 20), ("300000", 300000), ("1KB", 256 << 10), ("3MB", 3 << 20)])
        ct = rng.choice([None, None, "text", "bin", "audio", "video"])
        gpus = rng.randint(1, args.max_gpus)
        use_stdin = rng.random() < 0.25 and mode != "bpe"      # BPE over stdin has its own (documented) chunking
        cmd = [BLT]
        om = None
        if mode == "bpe":
            k = rng.choice([1, 50, 256, 5000, 40000])
            l, r = synth.merges_from_sample(data if n >= 2 else np.frombuffer(b"ab", np.uint8), k)
            mp = os.path.join(d, "m.txt")
            synth.write_merges_file(mp, l, r)
            om = ora.Merges.from_file(mp)
            cmd += ["--merges", mp]
        if mode == "passthrough":
            cmd += ["--passthrough"]
        if cs_arg: cmd += ["--chunksize", cs_arg]
        if ct: cmd += ["--type", ct]
        cmd += ["--gpus", str(gpus)]
        want = bytes(ora.run_buffer(mode, data, eff, 4, om, TYPES[ct] if ct else None))
        inp, outp = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        if use_stdin:
            r = subprocess.run(cmd, input=data.tobytes(), capture_output=True)
            got = r.stdout
        else:
            data.tofile(inp)
            r = subprocess.run(cmd + ["-i", inp, "-o", outp], capture_output=True)
            got = open(outp, "rb").read() if os.path.exists(outp) else b""
        cases += 1
        if r.returncode != 0 or got != want:
            fails += 1
            print("MISMATCH", dict(n=n, mode=mode, chunk=cs_arg, type=ct, gpus=gpus, stdin=use_stdin, rc=r.returncode,
                                   got=len(got), want=len(want), err=r.stderr[-200:]), flush=True)
        # and back again
        if mode != "passthrough" and not use_stdin and r.returncode == 0 and got == want:
            back = os.path.join(d, "back.bin")
            cmd2 = [BLT, "--detokenize", "-i", outp, "-o", back] + (["--merges", mp] if mode == "bpe" else []) + (["--type", ct] if ct else [])
            r2 = subprocess.run(cmd2, capture_output=True)
            if r2.returncode != 0 or open(back, "rb").read() != data.tobytes():
                fails += 1
                print("ROUND TRIP MISMATCH", dict(n=n, mode=mode, rc=r2.returncode, err=r2.stderr[-200:]), flush=True)
print(f"cli fuzz: {cases} cases, {fails} failures, seed {args.seed}")
sys.exit(1 if fails else 0)
