#!/usr/bin/env python
"""One cloud node serving B edge drones per planning step, end to end:

  signed state messages in  ->  BatchedMissionGoals (goal per drone)  ->  ONE batched solve through
  BatchWorkspace(outputs="solution").solve_host (the kernel reads pinned host inputs and writes
  one result row per drone back over PCIe)  ->  one signed trajectory message per drone out.

    python examples/serve_drones.py [B] [steps]

The wire format is the reference's (communication/secure_serializer.py:92-169), so an unmodified
edge client verifies and decodes the answers.  Needs a CUDA device."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dart_planner_b200 as dp  # noqa: E402
from dart_planner_b200.config import make_params  # noqa: E402
from dart_planner_b200.mission import BatchedMissionGoals, SemanticWaypoint  # noqa: E402
from dart_planner_b200.planner import BatchWorkspace  # noqa: E402
from dart_planner_b200.wire import SignedEnvelope, batch_trajectory_payloads, trajectory_from_payload  # noqa: E402


def serve(B=256, steps=5, secret="demo-secret", verbose=True):
    cfg = dp.SE3MPCConfig.from_yaml() if hasattr(dp.SE3MPCConfig, "from_yaml") else dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)
    params = make_params(cfg)
    N, dt = int(params.horizon), float(params.dt)
    cloud, edge = SignedEnvelope(secret_key=secret), SignedEnvelope(secret_key=secret)
    goals = BatchedMissionGoals(B, [SemanticWaypoint([10.0, 0.0, 5.0], "safe_zone"),
                                    SemanticWaypoint([15.0, 10.0, 8.0], "observation_point")])
    ws = BatchWorkspace(params, B, pinned=True, outputs="solution")
    rng = np.random.default_rng(0)
    p = rng.uniform(-3, 3, (B, 3))
    p[:, 2] = rng.uniform(1.0, 3.0, B)
    v = np.zeros((B, 3))
    stats = []
    for k in range(steps):
        now = 100.0 + k * dt
        # edge -> cloud: one signed state message per drone (what zmq_client sends)
        inbox = [edge.serialize({"command": "state", "drone_id": b, "position": p[b], "velocity": v[b]}) for b in range(B)]
        t0 = time.perf_counter()
        states = [cloud.deserialize(m) for m in inbox]
        P = np.array([s["position"] for s in states])
        V = np.array([s["velocity"] for s in states])
        G = goals.get_current_goals(P, now)
        t1 = time.perf_counter()
        sol = ws.solve_host(P, V, G)                       # ONE launch for all drones
        t2 = time.perf_counter()
        outbox = [cloud.serialize(d) for d in batch_trajectory_payloads(sol, now, dt, ids=list(range(B)))]
        t3 = time.perf_counter()
        # cloud -> edge: every drone verifies its answer and flies the first step of it
        for b, m in enumerate(outbox):
            tr = trajectory_from_payload(edge.deserialize(m))
            assert tr.positions.shape == (N, 3)
        T0 = sol.thrust_vectors[:, 0]
        a = T0 / float(params.mass) - np.array([0.0, 0.0, float(params.gravity)])
        p, v = P + V * dt + 0.5 * a * dt * dt, V + a * dt
        stats.append((t1 - t0, t2 - t1, t3 - t2))
        if verbose:
            print(f"step {k}: decode+goals {1e3 * (t1 - t0):.2f} ms | batched solve {1e3 * (t2 - t1):.3f} ms "
                  f"({B / (t2 - t1) / 1e6:.2f} M solves/s host to host) | sign {1e3 * (t3 - t2):.2f} ms | "
                  f"phases {np.bincount(goals.phase, minlength=6).tolist()}")
    return sol, outbox, stats


if __name__ == "__main__":
    serve(int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 5)
