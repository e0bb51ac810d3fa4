#!/usr/bin/env python
"""Executed-instruction mix of one kernel from an .ncu-rep captured with --set full --import-source on (source page, SASS
view): warp-instructions executed and shared-memory wavefronts per opcode, per unit of work.
    python tools/ncu_instr_mix.py file.ncu-rep <units> [title]        (units: e.g. input bytes / 512 = warp-rounds)"""
import collections, csv, io, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
title = sys.argv[3] if len(sys.argv) > 3 else rep
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = next(i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address")
hdr = rows[h]
i_src, i_exec, i_wf = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared")
ops, wfs = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    try:
        n, w = int(r[i_exec]), int(r[i_wf])
    except (ValueError, IndexError):
        continue
    tok = r[i_src].split()
    if not tok:
        continue
    op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
    ops[op] += n
    wfs[op] += w
print(title)
print("warp-instructions executed and shared-memory wavefronts per unit (opcode rows of the source page)")
for op, n in ops.most_common(40):
    print(f"{op:34s} {n / units:8.2f}   wavefronts {wfs[op] / units:7.2f}")
print(f"{'sum':34s} {sum(ops.values()) / units:8.2f}   wavefronts {sum(wfs.values()) / units:7.2f}")
