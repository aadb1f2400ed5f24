#!/usr/bin/env python
"""Reduce an `ncu --page raw --csv` export to the columns DESIGN.md cites: duration, tensor-pipe activity, issue-stall
reasons (per warp-cycle ratios), L2->SM bytes, L2 / DRAM throughput, registers, grid.  Usage: ncu_select.py raw.csv out.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
want += sorted(h for h in hdr if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio"))
want += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith(".ratio") and h not in want)
idx = [hdr.index(w) for w in want if w in hdr]
w = csv.writer(open(sys.argv[2], "w"))
w.writerow([hdr[i] for i in idx])
for r in data:
    row = [r[i] for i in idx]
    m = re.search(r"([A-Za-z0-9_]+<[^(]*>|[A-Za-z0-9_]+)\(", row[0].replace("(int)", ""))
    row[0] = m.group(1) if m else row[0][:60]
    w.writerow(row)
print("kept", len(idx), "columns,", len(data), "kernels")
