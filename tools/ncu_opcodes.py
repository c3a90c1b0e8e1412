#!/usr/bin/env python
"""Executed warp instructions per SASS opcode (and per pipe class) of one profiled kernel.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
       python tools/ncu_opcodes.py sass.csv [solves]
With `solves` (problems the launch solved) the counts are also given per solve."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
solves = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
i_src, i_ex, i_smp = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
i_thr = H.index("Thread Instructions Executed")
ops, smp = collections.Counter(), collections.Counter()
thr = 0
for r in rows[hdr + 1:]:
    if len(r) <= i_ex:
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src])
    if not m:
        continue
    op = m.group(1)
    n = int(r[i_ex] or 0)
    ops[op] += n
    smp[op] += int(r[i_smp] or 0)
    thr += int(r[i_thr] or 0)


def klass(op):
    b = op.split(".")[0]
    if b in ("DADD", "DMUL", "DFMA"):
        return "fp64 arithmetic (DADD/DMUL/DFMA)"
    if b in ("DSETP", "DMNMX"):
        return "fp64 compare (DSETP)"
    if b == "MUFU":
        return "MUFU (rcp64h / rsq64h seeds)"
    if b in ("SHFL", "VOTE", "VOTEU", "MATCH", "REDUX"):
        return "shuffles / votes"
    if b in ("FSEL", "SEL", "SELP"):
        return "selects (a double select is two)"
    if b in ("MOV", "IMAD", "UMOV", "PRMT", "CS2R", "S2R", "R2UR", "LDC", "LDCU", "ULDC", "S2UR"):
        return "moves / IMAD / constant loads" if not op.startswith("IMAD.WIDE") else "integer / address"
    if b in ("LDS", "STS", "LDSM"):
        return "shared memory"
    if b in ("LDL", "STL"):
        return "local memory (S/Y pairs, spills)"
    if b in ("LDG", "STG", "LD", "ST", "ATOMG", "ATOM", "RED", "MEMBAR", "ERRBAR", "CCTL"):
        return "global memory"
    if b in ("BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "BREAK", "WARPSYNC", "BAR", "NOP", "JMP", "BRX", "YIELD", "NANOSLEEP", "BMOV", "DEPBAR"):
        return "control (branches, barriers, syncs)"
    if b in ("ISETP", "PLOP3", "LOP3", "IADD3", "IADD", "LEA", "SHF", "POPC", "FLO", "IABS", "IMNMX", "VIADD", "VIMNMX", "ULOP3", "UIADD3", "USHF", "ULEA", "UISETP", "UIMAD", "USEL", "P2R", "R2P", "BREV", "I2F", "F2I", "I2FP", "F2F", "F2FP", "UPLOP3", "UFLO", "UPOPC", "ISCADD", "LOP"):
        return "integer / predicate logic"
    return "other"


total = sum(ops.values())
cl, cs = collections.Counter(), collections.Counter()
for op, n in ops.items():
    cl[klass(op)] += n
    cs[klass(op)] += smp[op]
tsm = sum(smp.values()) or 1
print(f"executed warp instructions {total}" + (f" = {total / solves:.0f} per solve" if solves else "") +
      f"; avg active threads per instruction {thr / total:.1f}")
print(f"{'class':48s} {'executed':>12s} {'share':>7s} {'samples':>8s}" + ("   per solve" if solves else ""))
for k, n in cl.most_common():
    print(f"{k:48s} {n:12d} {100 * n / total:6.1f}% {100 * cs[k] / tsm:7.1f}%" + (f" {n / solves:10.1f}" if solves else ""))
print()
print("top opcodes:")
for op, n in ops.most_common(28):
    print(f"  {op:28s} {n:12d} {100 * n / total:6.1f}%" + (f" {n / solves:10.1f}" if solves else ""))
