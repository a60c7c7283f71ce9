#!/usr/bin/env python
"""Per-source-line summary of one kernel of an ncu report (needs -lineinfo + --import-source on):
    python tools/ncu_lines.py report.ncu-rep <kernel regex> [top N]
prints instructions executed and stall samples per CUDA source line, heaviest first."""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, lines, tot_i, tot_s = None, [], 0, 0
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or not r[0].isdigit():
        continue
    off = len(r) - len(hdr)          # unescaped quotes / commas in the source text add columns
    try:
        ins = int(r[hdr["Instructions Executed"] + off] or 0)
        smp = int(r[hdr["# Samples"] + off] or 0)
    except ValueError:
        continue
    def col(name):
        try:
            return int(r[hdr[name] + off] or 0)
        except (ValueError, KeyError, IndexError):
            return 0
    st = {k: col(k) for k in ("stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_sleep", "stall_lg", "stall_mio", "stall_math")}
    lines.append((ins, smp, cur_file, int(r[0]), r[1].strip()[:100], st))
    tot_i += ins
    tot_s += smp
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
print("by instructions:")
for ins, smp, f, ln, src, st in sorted(lines, key=lambda t: -t[0])[:top]:
    print(f"{100.0 * ins / max(tot_i, 1):5.1f}% inst {100.0 * smp / max(tot_s, 1):5.1f}% smp  {f}:{ln}  {src}")
print("by stall samples:")
for ins, smp, f, ln, src, st in sorted(lines, key=lambda t: -t[1])[:top // 2]:
    why = " ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{100.0 * ins / max(tot_i, 1):5.1f}% inst {100.0 * smp / max(tot_s, 1):5.1f}% smp  {f}:{ln}  {src[:70]}  [{why}]")
tot = {}
for *_, st in lines:
    for k, v in st.items():
        tot[k] = tot.get(k, 0) + v
print("stall totals:", {k[6:]: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1])})
