#!/usr/bin/env python
"""BASELINE configs[3] under torchrun: 1 M Monte-Carlo initial-state solves (seed 3, one goal
(10,0,5), p0~N((0,0,2),1), v0~N(0,0.5^2)) sharded by problem index over the ranks, one NCCL
gather to rank 0, checked bit-for-bit against rank 0 solving the whole batch alone.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tools/shard_check.py [B]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dart_planner_b200 as dp  # noqa: E402
from dart_planner_b200.config import make_params  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(3)
p0 = rng.normal((0, 0, 2), 1.0, (B, 3))
v0 = rng.normal(0, 0.5, (B, 3))
goal = np.tile([10.0, 0.0, 5.0], (B, 1))
params = make_params(dp.SE3MPCConfig(prediction_horizon=8, dt=0.1))
solver = dp.ShardedSolver(params)
solver.solve(p0, v0, goal)                        # warm-up (allocations, NCCL communicator)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
sol = solver.solve(p0, v0, goal)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    # rank 0 alone through the same row path (the same kernel build: bit-identical), and through
    # plan_batch (the throughput build at this size: same counters, x to rounding -- the builds
    # contract a few products differently, tools/variant_diff.py)
    from dart_planner_b200.planner import BatchWorkspace, HostSolution
    ws = BatchWorkspace(params, B, pinned=False, outputs="all")
    ws.set_inputs_device(p0, v0, goal)
    alone = HostSolution.from_packed_rows(8, ws.solve_rows_device().cpu().numpy())
    ok = (np.array_equal(sol.x, alone.x) and np.array_equal(sol.cost, alone.cost)
          and np.array_equal(sol.nfev, alone.nfev) and np.array_equal(sol.status, alone.status)
          and np.array_equal(sol.body_rates, alone.body_rates))
    other = dp.plan_batch(p0, v0, goal, dp.SE3MPCConfig(prediction_horizon=8, dt=0.1), to_host=True)
    cross = float(np.abs(sol.x - other.x).max())
    same_counters = bool(np.array_equal(sol.nit, other.nit) and np.array_equal(sol.nfev, other.nfev)
                         and np.array_equal(sol.status, other.status))
    print(f"shard_check world={world} B={B} identical_to_single_gpu={ok} nit_hist={np.bincount(sol.nit).tolist()} "
          f"sharded solve {dt * 1e3:.1f} ms host arrays in -> HostSolution out (stages: {solver.last_timing}); "
          f"vs the throughput build: max |dx| {cross:.2e}, counters equal {same_counters}", flush=True)
    assert ok and same_counters and cross < 1e-12
dist.barrier()
dist.destroy_process_group()
