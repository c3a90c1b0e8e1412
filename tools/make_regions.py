#!/usr/bin/env python
"""Write a regions file (lo hi name) for tools/ncu_regions.py from markers in se3mpc_core.cuh."""
import re
import sys

src = open("dart_planner_b200/csrc/se3mpc_core.cuh").read().split("\n")
marks = [
    (r"^DP_HD double DP_MUL", "rounding/division/min-max helpers"), (r"^struct SeqGroup", "lane-group helpers (shuffles)"),
    (r"^struct LineSearch", "dcstep"), (r"^DP_HD int dcsrch", "dcsrch"), (r"^DP_HD int chol_ut", "chol_ut"),
    (r"^DP_HD int trsl_ut", "trsl_ut"), (r"^struct GridPenalty", "grid penalty"), (r"^struct SolveStats", "solver: members/ctor"),
    (r"DP_HD double grad_at", "grad_at/gat/gold"), (r"DP_HD double eval_fg", "eval_fg"), (r"DP_HD double projgr", "projgr"),
    (r"DP_HD int ring\(", "ring/bmv"), (r"DP_HD int cauchy_prepare", "cauchy classify+breakpoints+closed form"),
    (r"DP_HD int cauchy_walk", "cauchy walk (stored pairs)"), (r"DP_HD bool is_free", "formk"),
    (r"DP_HD int cmprlb", "cmprlb"), (r"DP_HD int subsm", "subsm"), (r"DP_HD int update_memory", "update_memory"),
    (r"DP_HD void begin", "begin (clip, first evaluation)"), (r"DP_HD void iterate", "iterate: cauchy/subspace calls"),
    (r"---- lnsrlb ----", "iterate: lnsrlb setup (d, dtd, stpmx)"), (r"while \(!ls_done\)", "iterate: line search loop"),
    (r"/\* NEW_X \*/", "iterate: NEW_X tests + pair update"), (r"DP_HD void finish", "finish"),
    (r"DP_HD void cold_start", "cold/warm start"), (r"DP_HD void extract", "extract (SO(3))"),
]
pos = []
for pat, name in marks:
    for i, l in enumerate(src):
        if re.search(pat, l):
            pos.append((i + 1, name))
            break
    else:
        sys.exit(f"marker not found: {pat}")
pos.sort()
for (lo, name), (hi, _) in zip(pos, pos[1:] + [(len(src) + 1, "")]):
    print(lo, hi, name)
