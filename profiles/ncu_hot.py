#!/usr/bin/env python
"""Hot spots of one kernel of an .ncu-rep (source page): top stall-sample instructions, samples and executed
instructions per 200-instruction bucket, stall-reason mix of an index range.  Usage: ncu_hot.py rep [lo hi]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
iS, iE, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
d = rows[2:]
data = [(int(r[iS] or 0), int(r[iE] or 0), k, r[iSrc].strip()) for k, r in enumerate(d) if len(r) > iE]
tot = sum(x[0] for x in data)
print("total samples", tot, "instructions", len(data), "executed", sum(x[1] for x in data))
for s, e, k, t in sorted(data, reverse=True)[:16]:
    print("%6d %5.1f%% exec=%9d idx=%5d %s" % (s, 100 * s / tot, e, k, t[:90]))
b, ex = collections.Counter(), collections.Counter()
for s, e, k, t in data:
    b[k // 200] += s
    ex[k // 200] += e
print("buckets (idx, samples, executed):", [(k * 200, b[k], ex[k]) for k in sorted(b) if ex[k] > 0])
if len(sys.argv) > 3:
    lo, hi = int(sys.argv[2]), int(sys.argv[3])
    cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    c = collections.Counter()
    for r in d[lo:hi]:
        for col in cols:
            c[col] += int(r[hdr.index(col)] or 0)
    t = sum(c.values())
    print("stalls [%d, %d):" % (lo, hi), " ".join("%s:%.0f%%" % (k[6:], 100 * v / t) for k, v in c.most_common(9)))
    for k in range(lo, hi):
        r = d[k]
        if int(r[iS] or 0) > 0.002 * tot or "SYNCS" in r[iSrc] or "BAR" in r[iSrc]:
            print(k, r[iS], r[iE], r[iSrc].strip()[:100])
