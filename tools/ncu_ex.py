#!/usr/bin/env python
"""Per-source-line executed warp-instruction counts (top N) from an ncu report. usage: ncu_ex.py rep kernel cubin [launch] [N]"""
import csv, io, re, subprocess, sys, collections
rep, kname, cubin = sys.argv[1:4]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
line_of, cur, infn = {}, None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        infn = kname in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur: line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
blk = rows[starts[which]:starts[which + 1]]
hdr = blk[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in blk[2:] if len(r) == len(hdr)]
base = int(data[0][ix["Address"]], 16)
agg = collections.Counter(); smp = collections.Counter(); tot = 0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    ex = int(r[ix["Instructions Executed"]] or 0)
    key = line_of.get(off, ("?", 0))
    agg[key] += ex; smp[key] += int(r[ix["# Samples"]] or 0); tot += ex
print("total executed warp instructions", tot)
import glob
srcs = {}
for (f, l), ex in agg.most_common(topn):
    if f not in srcs:
        p = glob.glob(f"/root/repo/scalable-e3-gnn_b200/csrc/{f}")
        srcs[f] = open(p[0]).read().splitlines() if p else []
    text = srcs[f][l - 1].strip()[:80] if 0 < l <= len(srcs[f]) else ""
    print(f"{f}:{l:4d} {100.0*ex/tot:5.1f}% ex={ex:10d} smp={smp[(f,l)]:6d} | {text}")
