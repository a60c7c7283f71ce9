#!/usr/bin/env python
"""Per-CUDA-source-line stall samples: joins `ncu --page source --csv` (SASS view) with `nvdisasm -g` line info.
usage: ncu_lines.py rep.ncu-rep <kernel substring> <cubin> [launch index] [min %]"""
import csv, io, re, subprocess, sys, collections

rep, kname, cubin = sys.argv[1:4]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
minpct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
# offset -> (file, line) for the wanted function
line_of, cur, infn = {}, None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
blk = rows[starts[which]:starts[which + 1]]
hdr = blk[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in blk[2:] if len(r) == len(hdr)]
base = int(data[0][ix["Address"]], 16)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    s = int(r[ix["# Samples"]] or 0); ex = int(r[ix["Instructions Executed"]] or 0)
    key = line_of.get(off, (("?", 0), ""))[0]
    a = agg[key]; a[0] += s; a[1] += ex; tot += s
    for h in stalls:
        v = int(r[ix[h]] or 0)
        if v: a[2][h[6:]] += v
srcs = {}
print(blk[0][1], "total samples", tot)
for key, (s, ex, st) in sorted(agg.items(), key=lambda kv: kv[0]):
    if s * 100.0 / max(tot, 1) < minpct:
        continue
    f, l = key
    if f not in srcs:
        try:
            import glob
            p = glob.glob(f"/root/repo/scalable-e3-gnn_b200/csrc/{f}")
            srcs[f] = open(p[0]).read().splitlines() if p else []
        except Exception:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:70] if 0 < l <= len(srcs[f]) else ""
    top = ",".join(f"{k}:{v}" for k, v in st.most_common(2))
    print(f"{f}:{l:4d} {100.0*s/tot:5.1f}% ex={ex:9d} {top:34s} | {text}")
