#!/usr/bin/env python
"""Latency build vs throughput build vs row output on one large batch: counts the problems whose
results differ between the builds (they share the arithmetic; the exact-size code copies of the
throughput build may contract a product differently), and where.
usage: python tools/variant_diff.py [B] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dart_planner_b200 as dp  # noqa: E402
from dart_planner_b200.config import make_params  # noqa: E402
from dart_planner_b200.planner import BatchWorkspace, HostSolution  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(seed)
p0 = rng.uniform(-10, 10, (B, 3))
v0 = np.zeros((B, 3))
goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
params = make_params(dp.SE3MPCConfig(prediction_horizon=8, dt=0.1))
res = {}
for name, var in (("latency", "1"), ("throughput", "5")):
    os.environ["DART_SE3MPC_VARIANT"] = var
    ws = BatchWorkspace(params, B, pinned=False)
    ws.set_inputs_device(p0, v0, goal)
    sol = ws.solve_device()
    torch.cuda.synchronize()
    res[name] = (sol.out[:, :B].clone(), sol.meta[:, :B].clone())
    del ws
os.environ.pop("DART_SE3MPC_VARIANT")
wr = BatchWorkspace(params, B, pinned=True, outputs="solution")
wr.stage_host_inputs(p0, v0, goal)
rows = HostSolution.from_solution_rows(8, wr.solve_rows().numpy(), params)
a, b = res["latency"], res["throughput"]
dx = (a[0][:72] != b[0][:72]).any(dim=0)
dm = (a[1] != b[1]).any(dim=0)
print(f"B={B} seed={seed}: latency vs throughput build: {int(dx.sum())} problems differ in x, {int(dm.sum())} in counters; "
      f"max |dx| = {float((a[0][:72] - b[0][:72]).abs().max()):.3e}")
idx = torch.nonzero(dx)[:8, 0].tolist()
for i in idx:
    print("  problem", i, "nit/nfev/status latency", a[1][:3, i].tolist(), "throughput", b[1][:3, i].tolist(),
          "max|dx|", float((a[0][:72, i] - b[0][:72, i]).abs().max()))
xr = torch.as_tensor(rows.x).cuda().t()
dr = (xr != a[0][:72]).any(dim=0)
print(f"rows (latency build, zero-copy) vs latency SoA: {int(dr.sum())} problems differ; first {torch.nonzero(dr)[:8, 0].tolist()}")
