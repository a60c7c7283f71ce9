#!/usr/bin/env python
"""Compact per-kernel summary of an ncu report (one line per profiled launch):
    python tools/ncu_summary.py report.ncu-rep > profiles/<name>.csv"""
import csv
import subprocess
import sys

WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"),
        ("dram__bytes_write.sum", "dram_write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps_per_cycle"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "pipe_tensor_pct"),
        ("smsp__inst_executed.sum", "warp_instructions"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
w = csv.writer(sys.stdout)
cols = [(h, n) for h, n in WANT if h in idx]
stall = [(i, h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
         if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
w.writerow([n + (f" [{units[idx[h]]}]" if units[idx[h]] else "") for h, n in cols] + ["top stall reasons (share of samples)"])
for r in rows[2:]:
    st = []
    for i, h in stall:
        try:
            st.append((float(r[i]), h))
        except ValueError:
            pass
    tot = sum(v for v, _ in st) or 1.0
    top = "; ".join(f"{h} {100 * v / tot:.0f}%" for v, h in sorted(st, reverse=True)[:5])
    w.writerow([r[idx[h]][:80] for h, _ in cols] + [top])
