"""Counts the Blackwell-specific SASS instructions per kernel of the built library (no GPU needed):
UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA tensor loads/stores, UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, ACQBULK = griddepcontrol / bulk-group waits, REDG/ATOMG = global reductions.

    python tools/sass_summary.py [path/to/libuda_b200.so] > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "uda_aerial_semantic_segmentation_research_b200", "libuda_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTMALDG|UTMASTG|LDTM|STTM|UBLKCP|UTCBAR|SYNCS|ACQBULK|REDG|ATOMG)\b")
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
    elif cur:
        for t in pat.findall(line):
            cnt[cur][t] += 1
names = list(cnt)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
agg, inst = collections.defaultdict(collections.Counter), collections.Counter()
for k, d in zip(names, dem):
    d = d.replace("(anonymous namespace)::", "").replace("uda::", "").replace("tcconv::", "").replace("void ", "")
    base = d.split("<")[0].split("(")[0]
    agg[base] += cnt[k]
    inst[base] += 1
tot = collections.Counter()
for c in agg.values():
    tot += c
print(f"{os.path.basename(lib)}: Blackwell-specific SASS instructions (all template instances of a kernel summed)")
print("total: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items())))
for b, c in sorted(agg.items(), key=lambda kv: -(kv[1].get("UTCHMMA", 0) * 1000 + kv[1].get("UBLKCP", 0) + kv[1].get("UTMALDG", 0))):
    if c.get("UTCHMMA") or c.get("UBLKCP") or c.get("UTMALDG"):
        print(f"{b} [{inst[b]} instances]: " + ", ".join(f"{k} {v}" for k, v in sorted(c.items())))
