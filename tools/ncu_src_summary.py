#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: stall samples per region between marker instructions.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_src_summary.py src.csv [kernel_index]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# split per kernel ("Kernel Name" rows)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
blk = rows[starts[which]:starts[which + 1]]
print(blk[0][:2])
hdr = blk[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in blk[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
mark = re.compile(r"LDGSTS|UTC|LDTM|SYNCS|BAR\.|ATOM|RED\.|EXIT|FENCE|MEMBAR")
thr = int(sys.argv[3]) if len(sys.argv) > 3 else max(100, tot // 400)
acc = 0
for n, r in enumerate(data):
    s = int(r[ix["# Samples"]] or 0)
    acc += s
    src = r[ix["Source"]]
    if s > thr or mark.search(src):
        st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"{n:5d} {r[ix['Address']][-5:]} smp={s:6d} ex={r[ix['Instructions Executed']]:>9s} cum={100*acc/tot:5.1f}% {src[:64]:64s} {st[0] if s else ''}")
print("total samples", tot)
