#!/usr/bin/env python
"""Key counters of one kernel from an .ncu-rep (ncu --page raw --csv): duration, DRAM bytes, issue utilisation, pipe
utilisation, shared-memory wavefronts and bank conflicts, stall reasons per issued instruction.
    python tools/ncu_keys.py file.ncu-rep [kernel-substring]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, u = rows[0], rows[1]
pats = [r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^smsp__inst_executed\.sum$", r"^smsp__issue_active\.avg\.pct",
        r"^sm__inst_executed_pipe_(alu|fma|lsu|xu|uniform|adu|cbu)\.avg\.pct_of_peak_sustained_active$",
        r"^l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed$", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared(_op_ld|_op_st)?\.sum$",
        r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_ld|_op_st)?\.sum$", r"^smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio$",
        r"^smsp__warps_active\.avg\.per_cycle_active$", r"^sm__cycles_elapsed\.max$", r"^lts__t_sector_hit_rate\.pct$", r"^sm__throughput\.avg\.pct",
        r"^launch__registers_per_thread$", r"^smsp__cycles_active\.avg$"]
ki = h.index("Kernel Name") if "Kernel Name" in h else None
for r in rows[2:]:
    if ki is not None and sub not in r[ki]:
        continue
    print("==", r[ki][:100] if ki is not None else "")
    for a, b, c in zip(h, u, r):
        if any(re.search(p, a) for p in pats):
            try:
                v = float(c.replace(",", ""))
                if "stalled" in a and v < 0.05:
                    continue
                print(f"  {a:88s} {b:12s} {v:,.3f}")
            except ValueError:
                pass
