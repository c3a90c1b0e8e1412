#!/usr/bin/env python
"""Closed loop (BASELINE configs[4] inputs), 100 replans: time per sub-population count
(`ClosedLoopSim.run(parts=...)`) for the population sizes one GPU holds at 1 / 2 / 4 / 8 GPUs.
usage: python tools/closed_loop_parts.py [B list]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dart_planner_b200 as dp  # noqa: E402
from dart_planner_b200.config import make_params  # noqa: E402

Bs = [int(b) for b in sys.argv[1].split(",")] if len(sys.argv) > 1 else [65536, 32768, 16384, 8192]
params = make_params(dp.SE3MPCConfig(prediction_horizon=8, dt=0.1))
for B in Bs:
    rng = np.random.default_rng(4)
    p0 = rng.uniform(-10, 10, (B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    v0 = np.random.default_rng(44).uniform(-2, 2, (B, 3))
    sim = dp.ClosedLoopSim(params, B, plant_dt=0.1)
    out = []
    for parts in (1, 2, 3, 4, 6, 8):
        best = 1e9
        for _ in range(3):
            sim.reset(p0, v0, goal)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            sim.run(100, track_counters=False, parts=parts)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        out.append(f"{parts}: {best:6.2f} ms")
    print(f"B={B:6d} default parts {sim.default_parts()} | " + "  ".join(out), flush=True)
