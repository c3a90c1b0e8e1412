#!/usr/bin/env python
"""Workload for ncu: BASELINE configs[2] -- 65536 solves on a 256^3 grid with 64 sphere obstacles,
fused safety check; `penalty` as first argument adds the occupancy-grid penalty inside the solve
(gradient_mode 2: eight corner gathers per position and evaluation).
usage: python tools/profile_penalty.py [plain|penalty] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dart_planner_b200 as dp  # noqa: E402
from dart_planner_b200.config import make_params  # noqa: E402
from dart_planner_b200.planner import BatchWorkspace  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "penalty"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = 65536
rng = np.random.default_rng(2)
grid = dp.DenseOccupancyGrid((256, 256, 256), (-128, -128, -128), 0.2)
grid.add_obstacles(rng.uniform(-20, 20, (64, 3)), rng.uniform(0.5, 2.0, 64))
r2 = np.random.default_rng(2)
p0 = r2.uniform(-10, 10, (B, 3))
goal = np.concatenate([r2.uniform(-15, 15, (B, 2)), r2.uniform(3, 8, (B, 1))], axis=1)
v0 = np.random.default_rng(22).uniform(-2, 2, (B, 3))
cfg = dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)
ws = BatchWorkspace(make_params(cfg, gradient_mode=2 if mode == "penalty" else 0), B, pinned=False)
ws.set_inputs_device(p0, v0, goal)
ws.set_map(grid, 1.5, 0.6)
for _ in range(reps):
    ws.solve_device()
torch.cuda.synchronize()
print("done", mode, float(ws.out[72, :B].sum()), float((ws.hit[:B] >= 0).float().mean()))
