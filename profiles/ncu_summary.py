#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per kernel time, DRAM bytes, issue
utilisation, top stall reasons, top SASS opcodes.  Usage: ncu_summary.py file.ncu-rep"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1]
    print("==", name)
    for k in want:
        if k in hdr:
            print("   %-58s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    st = sorted(((float(r[hdr.index(s)] or 0), s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for s in stalls), reverse=True)
    print("   stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in st[:6]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, hdr2, agg = None, None, {}
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name":
        kern = r[1].split("(")[0].split("::")[-1]
        agg.setdefault(kern, collections.Counter())
        hdr2 = None
    elif r and r[0] == "Address":
        hdr2 = r
    elif hdr2 and kern and len(r) > 6:
        try:
            n = int(r[5])
        except ValueError:
            continue
        tok = r[1].split()
        op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")
        agg[kern][op.split(".")[0]] += n
for k, c in agg.items():
    t = sum(c.values())
    print("== opcodes", k, t)
    print("   " + " ".join("%s:%.1f%%" % (o, 100.0 * n / t) for o, n in c.most_common(14)))
