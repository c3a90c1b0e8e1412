#!/usr/bin/env python
"""Small mixed workload touching every kernel and mode once (handy under a debugger or a memory
checker where one is available): cold + warm solves at several
horizons, the fused map check, the penalty mode, a closed-loop step and the mapper kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dart_planner_b200 as dp  # noqa: E402
from dart_planner_b200.config import make_params  # noqa: E402

rng = np.random.default_rng(0)
grid = dp.DenseOccupancyGrid((64, 64, 64), (-32, -32, -32), 0.5)
grid.add_obstacles(rng.uniform(-10, 10, (8, 3)), rng.uniform(0.5, 2.0, 8))
for N, B in ((8, 97), (6, 33), (13, 21), (40, 9), (3, 50)):
    p0 = rng.uniform(-10, 10, (B, 3)); v0 = rng.uniform(-2, 2, (B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    goal[::5] = p0[::5] + 0.001
    cfg = dp.SE3MPCConfig(prediction_horizon=N, dt=0.1)
    a = dp.plan_batch(p0, v0, goal, cfg, grid=grid, safety_margin=1.0, to_host=True)
    xw = a.x.copy(); xw[::2, 6 * N:] += rng.normal(0, 0.5, xw[::2, 6 * N:].shape)
    dp.plan_batch(p0, v0, goal, cfg, x_warm=xw, to_host=True)
    dp.plan_batch(p0, v0, goal, cfg, gradient_mode=1, to_host=True)
    dp.plan_batch(p0, v0, goal, cfg, grid=grid, obstacle_penalty=True, to_host=True)
sim = dp.ClosedLoopSim(make_params(dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)), 70, plant_dt=0.1)
sim.reset(p0[:70] if len(p0) >= 70 else np.tile(p0, (2, 1))[:70], np.zeros((70, 3)), np.tile([1.0, 2.0, 5.0], (70, 1)))
sim.run(3)
grid.update_map(rng.uniform(-5, 5, (200, 3)), rng.normal(0, 1, (200, 3)), rng.uniform(0.5, 20, 200), 10.0)
grid.trace_rays(rng.uniform(-5, 5, (50, 3)), rng.normal(0, 1, (50, 3)), rng.uniform(0.5, 10, 50), max_vox=64)
torch.cuda.synchronize()
print("mixed_case done")
