#!/usr/bin/env python
"""Static SASS statistics of the solve-kernel instantiations in one object file / library:
instruction count, moves, selects per kernel.  usage: python tools/sass_static.py file.o [name-substring]"""
import collections
import re
import subprocess
import sys

out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else "se3mpc_solve_kernel"
cur = None
stats = collections.defaultdict(collections.Counter)
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and cur and want in cur:
        op = m.group(1)
        c = stats[cur]
        c["all"] += 1
        b = op.split(".")[0]
        if op.startswith("IMAD.MOV") or b == "MOV":
            c["mov"] += 1
        if b in ("FSEL", "SEL"):
            c["sel"] += 1
        if b in ("DADD", "DMUL", "DFMA"):
            c["f64"] += 1
        if b in ("BRA", "BSSY", "BSYNC"):
            c["ctl"] += 1
for k, c in stats.items():
    print(f"{k[:110]:110s} all={c['all']:6d} mov={c['mov']:5d} sel={c['sel']:5d} f64={c['f64']:5d} ctl={c['ctl']:5d}")
