#!/usr/bin/env python
"""Phase timeline of ONE warp of the latency build (cycles between the DP_TICK marks of
se3mpc_core.cuh), on the bench distribution.

  python tools/phase_timing.py build     # here: compiles gpurun_scratch/libdart_phase.so with -DDART_PHASE_TIMING
  python tools/phase_timing.py run [B]   # on the GPU box: prints the timeline of problem 0's warp
"""
import ctypes as C
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_scratch")
LIB = os.path.join(OUT, "libdart_phase.so")
NAMES = {0: "start: load inputs, cold start", 1: "begin: clip", 2: "begin: first f, projected gradient",
         10: "iteration top", 11: "cauchy: status, direction, breakpoints (+ closed form)",
         12: "cauchy: breakpoint walk with stored pairs", 13: "formk", 14: "cmprlb", 15: "subsm",
         16: "lnsrlb: d, dtd, step bound", 17: "line search: trial point + f", 18: "line search: exit",
         19: "NEW_X: projected gradient, tests, y", 30: "  formk: pair sums", 31: "  subsm: W'd sums",
         32: "  subsm: K^-1 solves", 33: "  subsm: step + projection", 40: "pair update + finish",
         41: "epilogue: stores, SO(3) extraction", 49: "  cauchy: entry", 50: "  cauchy: per-variable pass",
         51: "  cauchy: breakpoint count (reduction)", 52: "  cauchy: closed-form pass",
         53: "  cauchy: crossing count (reduction)"}


def build():
    sys.path.insert(0, ROOT)
    from dart_planner_b200 import build as b
    os.makedirs(OUT, exist_ok=True)
    objs = []

    def one(unit):
        obj = os.path.join(OUT, unit.replace(".cu", ".o"))
        cmd = [b.nvcc()] + b.ARCH + b.COMMON + b.UNITS[unit] + ["-DDART_PHASE_TIMING", "-c", os.path.join(b.CSRC, unit), "-o", obj]
        subprocess.run(cmd, check=True, capture_output=True)
        return obj

    with ThreadPoolExecutor(8) as ex:
        objs = list(ex.map(one, b.UNITS))
    subprocess.run([b.nvcc()] + b.ARCH + ["-shared", "-o", LIB] + objs, check=True)
    for o in objs:
        os.remove(o)
    print(LIB)


def run(B):
    os.environ["DART_SE3MPC_LIB"] = LIB
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import dart_planner_b200 as dp
    from dart_planner_b200 import _cabi
    from dart_planner_b200.config import make_params
    from dart_planner_b200.planner import BatchWorkspace
    L = _cabi.lib()
    rng = np.random.default_rng(1)
    p0 = rng.uniform(-10, 10, (B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    ws = BatchWorkspace(make_params(dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)), B, pinned=False)
    ws.set_inputs_device(p0, np.zeros((B, 3)), goal)
    buf = (C.c_longlong * 4000)()
    for _ in range(3):
        ws.solve_device()
        torch.cuda.synchronize()
        n = L.dart_phase_log_read(buf, 2000)
    print(f"B={B}; the logged warp holds problems 0..3: nit {ws.meta[0, :4].cpu().numpy()}, nfev {ws.meta[1, :4].cpu().numpy()}")
    print(" cycles   (+delta)  phase that ENDS here -> next phase starts")
    t0, prev = buf[1], buf[1]
    for i in range(n):
        idv, c = buf[2 * i], buf[2 * i + 1]
        print(f"{c - t0:8d} (+{c - prev:6d})  {idv:2d} {NAMES.get(idv, '')}")
        prev = c


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 4096)
