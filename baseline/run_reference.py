#!/usr/bin/env python
"""Times the UNMODIFIED reference planner (DART-Planner `SE3MPCPlanner.plan_trajectory`,
src/dart_planner/planning/se3_mpc_planner.py:215-228) on host cores.  TEST / BENCH INFRASTRUCTURE:
the product never imports this.

The reference is pure Python; `__graft_entry__.build()` copies the 14 modules its planner imports
from /root/reference into the git-ignored baseline/_ref/ (they travel to the GPU box with the
snapshot, never into the history) next to the unit-transparent `pint` stand-in of tools/refshim
(pint is not installed).  Nothing here alters the reference's code path: dt and
the horizon are set through its own mechanisms (TimingConfig(control_frequency=1/dt),
SE3MPCConfig(prediction_horizon=N)), the 0.5 m goal hysteresis (:199) is reset between problems.

  python baseline/run_reference.py --problems 4096 --procs 16 [--horizon 8 --dt 0.1 --seed 1]

prints one JSON line: aggregate solves/s over `procs` processes (one per core, BLAS threads = 1),
per-process rates, single-solve p50/p95 of the G2 problem, and the G2 optimum as a self-check.
"""
import argparse
import json
import logging
import os
import sys
import time

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = "1"

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "dart_planner", "planning", "se3_mpc_planner.py"))


def _planner(N, dt):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    logging.disable(logging.CRITICAL)
    import warnings
    warnings.filterwarnings("ignore")
    from dart_planner.common import timing_alignment as ta
    from dart_planner.common.types import DroneState
    from dart_planner.planning import se3_mpc_planner as ref
    ta.reset_timing_manager()
    ta.get_timing_manager(ta.TimingConfig(control_frequency=1.0 / dt))
    pl = ref.SE3MPCPlanner(ref.SE3MPCConfig(prediction_horizon=N))
    assert abs(pl.se3_config.dt - dt) < 1e-15
    return pl, DroneState


def workload(B, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    p0 = rng.uniform(-10, 10, (B, 3))
    v0 = np.zeros((B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    return p0, v0, goal


def _worker(job):
    N, dt, seed, B, lo, hi, warm = job
    import numpy as np
    pl, DroneState = _planner(N, dt)
    p0, v0, goal = workload(B, seed)
    for b in range(lo, min(hi, lo + warm)):                     # warm-up solves (imports, caches)
        pl.goal_position = None
        pl.plan_trajectory(DroneState(timestamp=0.0, position=p0[b].copy(), velocity=v0[b].copy()), goal[b].copy())
    t0 = time.perf_counter()
    fsum = 0.0
    for b in range(lo, hi):
        pl.goal_position = None                                 # defeat the goal hysteresis (:199)
        tr = pl.plan_trajectory(DroneState(timestamp=0.0, position=p0[b].copy(), velocity=v0[b].copy()),
                                goal[b].copy())
        fsum += float(tr.positions[-1][0])
    return hi - lo, time.perf_counter() - t0, fsum


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=4096)
    ap.add_argument("--procs", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--horizon", type=int, default=8)
    ap.add_argument("--dt", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--warm", type=int, default=20)
    a = ap.parse_args()
    if not available():
        print(json.dumps({"unavailable": "baseline/_ref is empty (built from /root/reference by __graft_entry__.build())"}))
        return
    import multiprocessing as mp
    import numpy as np
    import statistics
    B, P = a.problems, max(1, a.procs)
    edges = [B * i // P for i in range(P + 1)]
    jobs = [(a.horizon, a.dt, a.seed, B, edges[i], edges[i + 1], a.warm) for i in range(P) if edges[i + 1] > edges[i]]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[1] for r in res)
    # single-solve latency of G2 (N=8, dt=0.1, p0=(0,0,2) -> goal (10,0,5); f* = 4370.579984226455)
    pl, DroneState = _planner(8, 0.1)
    lat = []
    for i in range(260):
        pl.goal_position = None
        st = DroneState(timestamp=0.0, position=np.array([0.0, 0.0, 2.0]), velocity=np.zeros(3))
        t1 = time.perf_counter()
        pl.plan_trajectory(st, np.array([10.0, 0.0, 5.0]))
        lat.append((time.perf_counter() - t1) * 1e3)
    lat = sorted(lat[60:])
    sol = pl._solve_se3_mpc(DroneState(timestamp=0.0, position=np.array([0.0, 0.0, 2.0]), velocity=np.zeros(3)))
    cpu = ""
    try:
        with open("/proc/cpuinfo") as fh:
            cpu = [l.split(":", 1)[1].strip() for l in fh if l.startswith("model name")][0]
    except Exception:
        pass
    import scipy
    print(json.dumps({
        "value": B / busy, "unit": "solves/s", "procs": len(jobs), "problems": B,
        "per_process_solves_per_s": [r[0] / r[1] for r in res], "slowest_process_s": busy, "pool_wall_s": wall,
        "single_solve_ms": {"p50": statistics.median(lat), "p95": lat[int(0.95 * len(lat))], "n": len(lat),
                            "problem": "G2: N=8, dt=0.1, (0,0,2)->(10,0,5)"},
        "g2_thrust_z": float(sol["thrust_vectors"][0][2]), "g2_expected_thrust_z": 14.5573723008,
        "cpu_model": cpu, "scipy": scipy.__version__, "numpy": np.__version__,
        "what": "unmodified reference SE3MPCPlanner.plan_trajectory, one process per core, BLAS threads 1, "
                "goal hysteresis reset per problem"}))


if __name__ == "__main__":
    main()
