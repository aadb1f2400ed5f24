#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per kernel launches, total time, share of the window, DRAM bytes and DRAM GB/s.
Usage: launch_shares.py launches.csv [out.csv] [--step]   (--step: only the last complete training step of the list)
(per-launch times are cold-cache and serialised: compare SHARES, not absolutes)."""
import collections
import csv
import re
import sys


def short(name):
    n = name.replace("void ", "")
    m = re.search(r"([A-Za-z0-9_]+)(<[^(]*>)?\(", n)
    return m.group(1) if m else n[:40]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, mi, vi, idi, ui = (h.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > vi:
            per.setdefault((int(r[idi]), r[ki]), {})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
    scale_t = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
    scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    # --step: keep only the LAST complete training step = the launches between the last two optimizer launches
    if "--step" in sys.argv:
        keys = list(per.keys())
        adam = [i for i, k in enumerate(keys) if "adam_kernel" in k[1]]
        if len(adam) >= 2:
            keep = set(k for k in keys[adam[-2] + 1:adam[-1] + 1] if "spin_kernel" not in k[1])   # torch.cuda._sleep
            per = collections.OrderedDict((k, v) for k, v in per.items() if k in keep)
    agg, tot = collections.OrderedDict(), 0.0
    for (_, name), m in per.items():
        t, u = m["gpu__time_duration.sum"]
        t *= scale_t.get(u, 1.0)
        rd = wr = 0.0
        if "dram__bytes_read.sum" in m:
            rd = m["dram__bytes_read.sum"][0] * scale_b[m["dram__bytes_read.sum"][1]]
            wr = m["dram__bytes_write.sum"][0] * scale_b[m["dram__bytes_write.sum"][1]]
        a = agg.setdefault(short(name), [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += rd; a[3] += wr
        tot += t
    lines = ["kernel,launches,total_us,share_pct,dram_read_mb,dram_write_mb,dram_gbs"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k},{a[0]},{a[1]:.1f},{100 * a[1] / tot:.2f},{a[2] / 1e6:.1f},{a[3] / 1e6:.1f},{(a[2] + a[3]) / a[1] / 1e3:.0f}")
    lines.append(f"TOTAL,{len(per)},{tot:.1f},100.00,,,")
    text = "\n".join(lines)
    print(text)
    outs = [a for a in sys.argv[2:] if not a.startswith("--")]
    if outs:
        open(outs[0], "w").write(text + "\n")


if __name__ == "__main__":
    main()
