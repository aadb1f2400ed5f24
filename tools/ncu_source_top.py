#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export (SASS view): for every profiled kernel the instructions with the
most warp-stall samples (+ one line of context), so that the stall site of each warp role (TMA producer, MMA issuer,
epilogue warps) is visible.  Usage: ncu_source_top.py file.csv[.gz] [top_n] [kernel-substring]"""
import csv, gzip, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 15; filt = sys.argv[3] if len(sys.argv) > 3 else ""
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
kernels, cur = [], None
for row in csv.reader(f):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}; kernels.append(cur); continue
    if row[0] == "Address":
        cur["hdr"] = row; continue
    if cur is not None and cur["hdr"] is not None:
        cur["rows"].append(row)
for ki, k in enumerate(kernels):
    if filt and filt not in k["name"]:
        continue
    h = {n: i for i, n in enumerate(k["hdr"])}
    si, ni, ii = h["# Samples"], h["Warp Stall Sampling (Not-issued Samples)"], h["Instructions Executed"]
    stall_cols = [(n, i) for n, i in h.items() if n.startswith("stall_")]
    rows = k["rows"]
    tot = sum(int(r[si] or 0) for r in rows)
    short = k["name"].split("::")[-1][:70]
    print(f"\n=== [{ki}] {short}   total samples {tot}")
    order = sorted(range(len(rows)), key=lambda j: -int(rows[j][si] or 0))[:top]
    for j in sorted(order):
        r = rows[j]
        s = int(r[si] or 0)
        st = sorted(((int(r[i] or 0), n) for n, i in stall_cols), reverse=True)[:2]
        ctx = rows[j - 1][1].strip()[:38] if j > 0 else ""
        print(f"  {100.0*s/max(tot,1):5.1f}%  #{j:5d} exec {r[ii]:>8s}  {r[1].strip()[:60]:60s} | prev: {ctx:38s} | {st}")
