#!/usr/bin/env python
"""Divergence-prone spots: BSSY (convergence-barrier) and BRA counts of one kernel per CUDA source line.
usage: nvdisasm -c -g <cubin> > dis.txt; python tools/sass_branches.py dis.txt <function-substring> [topN]"""
import collections
import re
import sys

path, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fn = cur = None
bssy, bra, tot = collections.Counter(), collections.Counter(), collections.Counter()
for l in open(path, errors="ignore"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if fn and want in fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        body = l.split("*/", 1)[1].split()
        op = body[1] if body[0].startswith("@") else body[0]
        op = op.split(".")[0]
        tot[cur] += 1
        if op == "BSSY":
            bssy[cur] += 1
        if op == "BRA":
            bra[cur] += 1
print(f"instructions {sum(tot.values())}  BSSY {sum(bssy.values())}  BRA {sum(bra.values())}")
src = {}
for (f, ln), c in sorted(bssy.items(), key=lambda kv: -kv[1])[:top]:
    if f not in src:
        try:
            src[f] = open("dart_planner_b200/csrc/" + f).read().splitlines()
        except OSError:
            src[f] = []
    text = src[f][ln - 1].strip() if 0 < ln <= len(src[f]) else ""
    print(f"{f}:{ln:5d}  BSSY {c:3d}  BRA {bra[(f, ln)]:3d}  insts {tot[(f, ln)]:4d}  {text[:90]}")
