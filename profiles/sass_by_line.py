#!/usr/bin/env python
"""Joins an ncu `--page source --csv` export with `nvdisasm -g` of the same cubin and prints
executed warp-instructions and stall samples per source line of one kernel.

usage: sass_by_line.py <ncu_source.csv> <nvdisasm_-g_output> <mangled-kernel-substring> [nwarps]
"""
import collections
import csv
import re
import sys

src_csv, sass, kern = sys.argv[1:4]
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break  # further launches of the same kernel follow; the first one is enough
    if len(r) > 10:
        data.append(r)
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
cur = None
maps = []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]+\*/", l):
        maps.append(cur)
assert len(maps) >= len(data), (len(maps), len(data))
ins = collections.Counter()
smp = collections.Counter()
for r, key in zip(data, maps):
    ins[key] += int(r[ix["Instructions Executed"]])
    smp[key] += int(r[ix["# Samples"]])
nw = float(sys.argv[4]) if len(sys.argv) > 4 else max(int(r[ix["Instructions Executed"]]) for r in data)
ts = sum(smp.values())
print(f"total {sum(ins.values()) / nw:.1f} instr/warp, {ts} samples")
for key, c in sorted(ins.items(), key=lambda kv: -kv[1])[:60]:
    print(f"{key[0]:>14s}:{key[1]:<5d} {c / nw:7.1f} instr/warp  {100 * smp[key] / ts:5.1f}% samples")
