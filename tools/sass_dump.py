#!/usr/bin/env python
"""SASS evidence for every kernel of libblt_cuda.so: profiles/<tag>_sass/<kernel>.sass.gz (address, predicate, opcode,
operands; the hex encodings are dropped) and profiles/<tag>_sass_summary.md (instruction mix per kernel and the
instructions that prove the data path: bulk copies, 16/32-byte loads and stores, shared-memory lookups, mbarrier waits).
    python tools/sass_dump.py [tag]"""
import collections, gzip, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
LIB = os.path.join(ROOT, "blt_b200", "lib", "libblt_cuda.so")
OUT = os.path.join(ROOT, "profiles", f"{TAG}_sass")
os.makedirs(OUT, exist_ok=True)
ALU = ("LOP3", "IADD3", "SHF", "PRMT", "ISETP", "SEL", "VIADD", "LEA", "MOV", "PLOP3", "VIMNMX", "IABS", "SGXT", "P2R", "R2P", "IADD", "VIADDMNMX")
FMA = ("IMAD", "HFMA2", "FFMA", "FMUL", "FADD")
EVIDENCE = ("UBLKCP", "LDG.E.128", "STG.E.128", "STG.E.ENL2.256", "LDS.U16", "LDS.128", "STS.U16", "STS.128", "ATOMS", "ATOMG", "RED.E",
            "SYNCS.PHASECHK.TRANS64.TRYWAIT", "SYNCS.ARRIVE.TRANS64", "REDUX", "VOTE.ANY", "SHFL", "LDG.E.64.STRONG.GPU", "STG.E.64.STRONG.GPU",
            "LDG.E.128.CONSTANT", "LDG.E.U16.CONSTANT")

txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", "-p", n], capture_output=True, text=True).stdout.strip()
kernels, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        kernels[cur].append((m.group(1), m.group(2).strip()))
rows = []
for name, ins in kernels.items():
    if "bltk" not in name or not ins:
        continue
    dm = demangle(re.sub(r"^__nv_static_\d+__[0-9a-f]+_\d+_kernels_cu_[0-9a-f]+_\d+_(?=_Z)", "", name))
    m2 = re.search(r"(\w+(?:<[^(]*>)?)\(", dm.replace("(anonymous namespace)::", "").replace("bltk::", ""))
    short = re.sub(r"[^A-Za-z0-9]+", "_", m2.group(1) if m2 else dm).strip("_")[:100]
    with gzip.open(os.path.join(OUT, short + ".sass.gz"), "wt") as f:
        f.write(f"// {name}\n")
        for addr, i in ins:
            f.write(f"{addr} {i}\n")
    ops = collections.Counter()
    for _, i in ins:
        if i.startswith("@"):
            i = i.split(None, 1)[1]
        ops[i.split()[0]] += 1
    cls = collections.Counter()
    for k, v in ops.items():
        b = k.split(".")[0]
        cls["alu" if b in ALU else "fma" if b in FMA else "smem" if b in ("LDS", "STS", "ATOMS", "LDSM") else
            "global" if b in ("LDG", "STG", "RED", "ATOMG", "UBLKCP", "LD", "ST") else "other"] += v
    ev = {e: sum(v for k, v in ops.items() if k.startswith(e)) for e in EVIDENCE}
    rows.append((short, len(ins), dict(cls), {k: v for k, v in ev.items() if v}))
with open(os.path.join(ROOT, "profiles", f"{TAG}_sass_summary.md"), "w") as f:
    f.write(f"# SASS summary ({TAG}): `cuobjdump -sass blt_b200/lib/libblt_cuda.so`, sm_100a\n\n"
            f"Full listings: `profiles/{TAG}_sass/<kernel>.sass.gz`.  Static instruction counts (the sweeps are unrolled straight-line code).\n\n")
    for short, n, cls, ev in rows:
        f.write(f"## {short}\n\n{n} instructions: " + ", ".join(f"{k} {v}" for k, v in sorted(cls.items(), key=lambda kv: -kv[1])) + "\n\n")
        f.write("evidence: " + (", ".join(f"`{k}` x{v}" for k, v in ev.items()) or "-") + "\n\n")
print(f"{len(rows)} kernels -> {OUT}")
