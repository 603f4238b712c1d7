#!/usr/bin/env python
"""Sums the per-instruction stall-reason samples of an ncu `--page source --csv` export
(first kernel in the file) and prints the share of each reason.
usage: stall_summary.py <ncu_source.csv>"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = {hdr[i]: 0 for i in cols}
ninst = 0
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) <= max(cols):
        continue
    ninst += int(r[hdr.index("Instructions Executed")] or 0)
    for i in cols:
        tot[hdr[i]] += int(r[i] or 0)
s = sum(tot.values())
print(f"warp-instructions executed: {ninst}; stall samples: {s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {k:28s} {100 * v / s:5.1f}%")
