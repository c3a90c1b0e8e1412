#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by CUDA source line.
usage: ncu -i rep --page source --csv --print-source sass,cuda > src.csv; python tools/ncu_lines.py src.csv [topN]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}


def num(s):
    try:
        return int(float(s))
    except ValueError:
        return 0


agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
text = {}
for r in rows[hi + 1:]:
    try:
        ln = int(r[0])
    except ValueError:
        continue
    text[ln] = r[1]
    a = agg[ln]
    a[0] += 1
    a[1] += num(r[ix["Instructions Executed"]])
    a[2] += num(r[ix["# Samples"]])
    a[3] += num(r[ix["stall_no_inst"]])
    a[4] += num(r[ix["Thread Instructions Executed"]])
tot = [sum(a[i] for a in agg.values()) or 1 for i in range(5)]
print(f"static sass {tot[0]}  executed {tot[1]}  samples {tot[2]}  no_inst samples {tot[3]}  avg threads/inst {tot[4]/tot[1]:.1f}")
for key, name in ((0, "static count"), (1, "executed"), (2, "stall samples")):
    print(f"--- top by {name}")
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
        print(f"{ln:5d} static={a[0]:5d} exec={a[1]/tot[1]*100:5.1f}% samp={a[2]/tot[2]*100:5.1f}% "
              f"thr={a[4]/max(a[1],1):4.1f} | {text[ln].strip()[:100]}")
