#!/usr/bin/env python
"""Runs of consecutive register moves in the static SASS of one kernel, with the source line they
are attributed to: the copies ptxas places in front of a branch for the other edge's register
assignment show up as runs of 10-25 moves.  usage: nvdisasm -c -g cubin > dis.txt;
python tools/sass_movruns.py dis.txt <function-substring> [min run]"""
import re
import sys

path, want = sys.argv[1], sys.argv[2]
minrun = int(sys.argv[3]) if len(sys.argv) > 3 else 8
cur_fn, cur = None, None
run, start = 0, None
tot = 0
out = []
for l in open(path, errors="ignore"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur_fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if not (cur_fn and want in cur_fn):
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if not m:
        continue
    op = m.group(2)
    if op.startswith("IMAD.MOV") or op == "MOV" or op == "CS2R":
        if run == 0:
            start = (m.group(1), cur)
        run += 1
    elif op in ("LDS.64", "LDS", "LDC.64", "LDC"):   # loads interleaved into a run do not end it
        continue
    else:
        if run >= minrun:
            out.append((run, start, op))
            tot += run
        run = 0
print(f"{len(out)} runs of >= {minrun} moves, {tot} moves in them")
for r, (addr, ln), nxt in out:
    print(f"  {r:3d} moves at {addr}  {ln}  followed by {nxt}")
