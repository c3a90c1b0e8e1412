#!/usr/bin/env python
"""Per CUDA source line of one file: executed warp instructions by opcode group.
usage: ncu -i rep --page source --csv --print-source sass,cuda > src.csv
       python tools/ncu_line_ops.py src.csv se3mpc_core.cuh [solves] [topN]
The dump has one section per source file; SASS rows follow the CUDA line they belong to."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
want = sys.argv[2]
solves = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
GROUPS = {"mov": ("IMAD.MOV", "MOV", "UMOV"), "sel": ("FSEL", "SEL"), "f64": ("DADD", "DMUL", "DFMA"),
          "dsetp": ("DSETP",), "shfl": ("SHFL",), "int": ("LOP3", "ISETP", "IADD3", "VIADD", "IMAD.IADD", "PLOP3", "LEA", "SHF")}
sec = None
agg = collections.defaultdict(collections.Counter)
text = {}
cur = None
for i, r in enumerate(rows):
    if r and r[0] == "File Path" or (len(r) >= 2 and r[0] in ("File", "Source File")):
        sec = r[1] if len(r) > 1 else None
    if r and r[0] == "Line No":
        # the section's file name is in the row(s) just above
        up = " ".join(" ".join(x) for x in rows[max(0, i - 2):i])
        sec = up
        H = r
        iex = H.index("Instructions Executed")
        continue
    if sec is None or want not in sec or len(r) < 8:
        continue
    if r[0].strip().isdigit():
        cur = int(r[0])
        text[cur] = r[1]
        continue
    if cur is None:
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
    if not m:
        continue
    op = m.group(1)
    try:
        n = int(r[iex])
    except ValueError:
        continue
    agg[cur]["all"] += n
    for g, pre in GROUPS.items():
        if any(op == p or op.startswith(p + ".") or op.startswith(p) and p.endswith("MOV") for p in pre):
            agg[cur][g] += n
tot = collections.Counter()
for a in agg.values():
    tot.update(a)
print("totals per solve: " + "  ".join(f"{k}={v / solves:.0f}" for k, v in tot.items()))
for key in ("all", "mov", "sel"):
    print(f"--- top lines by {key} (per solve)")
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
        print(f"{ln:5d} all={a['all'] / solves:6.1f} mov={a['mov'] / solves:5.1f} sel={a['sel'] / solves:5.1f} f64={a['f64'] / solves:5.1f} "
              f"dsetp={a['dsetp'] / solves:5.1f} shfl={a['shfl'] / solves:5.1f} int={a['int'] / solves:5.1f} | {text[ln].strip()[:90]}")
