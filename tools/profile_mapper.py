#!/usr/bin/env python
"""Fixed mapper workload for ncu (the bench's mapper legs): one LiDAR-like scan of 200 000 rays into
a 256^3 grid (update_map: ray walk with 64-bit hit/miss counters, then the Bayes apply pass), 4 Mi
occupancy queries, and the batched trajectory safety check on 65 536 x 8 positions."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dart_planner_b200 as dp  # noqa: E402

rng = np.random.default_rng(3)
R = 200_000
sensors = rng.uniform(-10, 10, (8, 3))
rpos = sensors[rng.integers(0, 8, R)]
rdir = rng.normal(0, 1, (R, 3))
rhit = rng.uniform(0.5, 30.0, R)
rhit[rng.random(R) < 0.25] = np.nan
g = dp.DenseOccupancyGrid((256, 256, 256), (-128, -128, -128), 0.2, max_range=25.0)
g.add_obstacles(rng.uniform(-20, 20, (64, 3)), rng.uniform(0.5, 2.0, 64))
for _ in range(2):
    upd = g.update_map(rpos, rdir, rhit, 30.0)
Q = 1 << 22
qpos = torch.as_tensor(rng.uniform(-25, 25, (Q, 3)), dtype=torch.float64, device="cuda")
for _ in range(2):
    occ = g.query_occupancy_batch(qpos)
B, N = 65536, 8
traj = torch.as_tensor(rng.uniform(-20, 20, (3 * N, B)), dtype=torch.float64, device="cuda")
for _ in range(2):
    hit = g.trajectories_safe_soa(traj, B, N, 1.5, 0.6)
torch.cuda.synchronize()
print("done", upd["updated_voxels"], float(occ.sum()), int((hit >= 0).sum()))
