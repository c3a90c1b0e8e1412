#!/usr/bin/env python
"""Small fixed workload for ncu: `python tools/profile_case.py [B] [reps] [N] [rows]` launches the
solve kernel `reps` times on the bench distribution (seed 1); with a fourth argument `rows` the
launches are the zero-copy end-to-end ones (pinned host inputs, result rows into pinned host memory)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dart_planner_b200 as dp  # noqa: E402
from dart_planner_b200.config import make_params  # noqa: E402
from dart_planner_b200.planner import BatchWorkspace  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = int(sys.argv[3]) if len(sys.argv) > 3 else 8
rng = np.random.default_rng(1)
p0 = rng.uniform(-10, 10, (B, 3))
goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
rows = len(sys.argv) > 4 and sys.argv[4] == "rows"
ws = BatchWorkspace(make_params(dp.SE3MPCConfig(prediction_horizon=N, dt=0.1)), B, pinned=rows)
ws.set_inputs_device(p0, np.zeros((B, 3)), goal)
if rows:
    ws.stage_host_inputs(p0, np.zeros((B, 3)), goal)
for _ in range(reps):
    if rows:
        ws.solve_rows()
    else:
        ws.solve_device()
torch.cuda.synchronize()
print("done", B, reps, float(ws.out[9 * N, :B].sum()))
