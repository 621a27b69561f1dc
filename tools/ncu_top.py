#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv`: total stall samples by reason and the hottest SASS lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = {h: 0 for h in stall_cols}
samples = 0
for r in data:
    samples += int(r[col["# Samples"]] or 0)
    for h in stall_cols:
        tot[h] += int(r[col[h]] or 0)
print("total samples", samples, " instructions", len(data))
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print("  %-26s %8d  %5.1f%%" % (h, v, 100.0 * v / max(samples, 1)))
print("hottest lines:")
for idx, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][col["# Samples"]] or 0))[:n]:
    top = sorted(((int(r[col[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
    print("%5d %7s exec=%9s  %-70s %s" % (idx, r[col["# Samples"]], r[col["Instructions Executed"]], r[col["Source"]][:70],
                                       " ".join("%s=%d" % (h[6:], v) for v, h in top if v)))
