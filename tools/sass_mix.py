#!/usr/bin/env python
"""Static instruction mix of one kernel in libblt_cuda.so (cuobjdump -sass): opcode histogram and the split
between the alu pipe (LOP3/IADD3/SHF/PRMT/ISETP/SEL/...), the fma pipe (IMAD*), shared-memory and control
instructions.  The fused sweep's tile body is straight-line code, so static counts / R approximate the
warp-instructions per 512-byte round.   python tools/sass_mix.py <substring of the mangled name> [--per N]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALU = ("LOP3", "IADD3", "SHF", "PRMT", "ISETP", "SEL", "VIADD", "LEA", "MOV", "PLOP3", "VIMNMX", "IABS", "SGXT", "P2R", "R2P", "IADD")
FMA = ("IMAD", "HFMA2", "FFMA", "FMUL", "FADD")
XU = ("POPC", "FLO", "BREV", "MUFU")


def main():
    pat = sys.argv[1]
    per = float(sys.argv[sys.argv.index("--per") + 1]) if "--per" in sys.argv else 1.0
    lib = os.path.join(ROOT, "blt_b200", "lib", "libblt_cuda.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, ops = None, collections.Counter()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or pat not in cur:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1).strip()
        if ins.startswith("@"):
            ins = ins.split(None, 1)[1]
        ops[ins.split()[0]] += 1
    tot = sum(ops.values())
    cls = collections.Counter()
    for k, v in ops.items():
        base = k.split(".")[0]
        if base in ALU: cls["alu"] += v
        elif base in FMA: cls["fma"] += v
        elif base in XU: cls["xu"] += v
        elif base in ("LDS", "STS", "LDSM", "ATOMS"): cls["smem"] += v
        elif base in ("LDG", "STG", "LD", "ST", "RED", "ATOMG", "UBLKCP", "LDC", "LDCU"): cls["mem"] += v
        elif base in ("BRA", "BSSY", "BSYNC", "BREAK", "EXIT", "WARPSYNC", "NANOSLEEP", "SYNCS", "BAR", "CALL", "RET", "YIELD"): cls["ctl"] += v
        else: cls["other"] += v
    print(f"total {tot}  per-unit {tot / per:.1f}")
    print({k: round(v / per, 1) for k, v in cls.most_common()})
    for k, v in ops.most_common(28):
        print(f"  {k:26s} {v:6d} {v / per:8.1f}")


if __name__ == "__main__":
    main()
