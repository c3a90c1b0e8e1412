#!/usr/bin/env python
"""Dynamic instruction / stall-sample share per region of se3mpc_core.cuh.
Joins `nvdisasm -c -g` (address -> source line; instructions inlined from CUDA headers are
attributed to the last line of our own sources seen before them) with
`ncu --page source --csv --print-source sass` (address -> executed count, samples).
usage: python tools/ncu_regions.py dis.txt sass.csv kernel-substring regions.txt
regions.txt: lines `lo hi name` over se3mpc_core.cuh line numbers."""
import collections
import csv
import re
import sys

dis, sasscsv, want, regfile = sys.argv[1:5]
regions = []
for l in open(regfile):
    p = l.split(None, 2)
    if len(p) == 3 and p[0].isdigit():
        regions.append((int(p[0]), int(p[1]), p[2].strip()))
line_at = {}
cur_fn, cur, active = None, ("?", 0), False
for l in open(dis, errors="ignore"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        active = want in m.group(1)
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f = m.group(1).split("/")[-1]
        if f.startswith("se3mpc"):
            cur = (f, int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*)", l)
    if m:
        line_at[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sasscsv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0, 0]
for r in rows[hi + 1:]:
    try:
        a = int(r[0], 16)
        e = int(r[ix["Instructions Executed"]])
    except ValueError:
        continue
    if base is None:
        base = a
    f, ln = line_at.get(a - base, ("?", 0))
    name = f"[{f}]"
    if f == "se3mpc_core.cuh":
        name = "core:other"
        for lo, hi_, n in regions:
            if lo <= ln < hi_:
                name = n
                break
    v = (1, e, int(r[ix["# Samples"]]), int(r[ix["Thread Instructions Executed"]]))
    for i in range(4):
        agg[name][i] += v[i]
        tot[i] += v[i]
print(f"static {tot[0]}  executed {tot[1]}  samples {tot[2]}  avg thr/inst {tot[3] / max(tot[1], 1):.1f}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:28s} static {a[0]:6d}  exec {a[1] / tot[1] * 100:5.1f}%  samples {a[2] / max(tot[2], 1) * 100:5.1f}%  thr/inst {a[3] / max(a[1], 1):5.1f}")
