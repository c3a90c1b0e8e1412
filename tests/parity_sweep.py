#!/usr/bin/env python
"""Extended randomised parity sweep, GPU kernel vs the oracle (test infrastructure, like tests/):
`python tests/parity_sweep.py [seeds]` runs seeds x 12 random configurations (horizons 2..40,
weights, bounds, dt, tolerance, iteration cap, up to 13 000 problems), cold and warm, and prints
per case the fraction within the north-star tolerance and the fraction with equal counters.
A case below the thresholds of tests/test_gpu_parity.py is re-examined: the oracle is run again
with every start position (and warm-start entry) moved by one ulp, and the case only counts as a FAIL if the kernel
agrees with the oracle worse than the oracle agrees with itself."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as oracle_mod
import dart_planner_b200 as dp
COST_RTOL, CTRL_ATOL = 1e-5, 1e-4
worst = dict(ok=1.0, same=1.0)
bad = 0
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    rng = np.random.default_rng(9000 + seed)
    for trial in range(12):
        N = int(rng.choice([2, 3, 4, 5, 6, 7, 8, 9, 12, 16, 17, 24, 32, 33, 40]))
        kw = dict(max_velocity=float(rng.uniform(3, 15)), max_thrust=float(rng.uniform(18, 40)),
                  min_thrust=float(rng.uniform(0.5, 4)), max_tilt_angle=float(rng.uniform(0.3, 1.2)),
                  position_weight=float(rng.uniform(10, 300)), velocity_weight=float(rng.uniform(1, 30)),
                  acceleration_weight=float(rng.uniform(0.2, 5)), thrust_weight=float(rng.uniform(0.02, 1)),
                  max_iterations=int(rng.integers(2, 30)), convergence_tolerance=float(rng.choice([0.1, 0.05, 0.01])))
        dt = float(rng.choice([0.0025, 0.05, 0.1, 0.2])); mass = float(rng.uniform(0.6, 3.0))
        B = int(rng.choice([97, 384, 5000, 13000]))
        p0 = rng.uniform(-10, 10, (B, 3)); v0 = rng.uniform(-3, 3, (B, 3))
        goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(2, 9, (B, 1))], axis=1)
        cfg = dp.SE3MPCConfig(prediction_horizon=N, dt=dt, **kw)
        op = oracle_mod.make_params(horizon=N, dt=dt, mass=mass, **kw)
        for mode in ("cold", "warm"):
            if mode == "cold":
                ref = oracle_mod.solve_batch(op, p0, v0, goal, nthreads=16)
                sol = dp.plan_batch(p0, v0, goal, cfg, mass=mass, to_host=True)
                xw = ref.x.copy(); xw[:, 6 * N:] += rng.normal(0, 0.3, xw[:, 6 * N:].shape)
            else:
                ref = oracle_mod.solve_batch(op, p0 + 0.1, v0, goal, x_warm=xw, nthreads=16)
                sol = dp.plan_batch(p0 + 0.1, v0, goal, cfg, mass=mass, x_warm=xw, to_host=True)
            relf = np.abs(sol.cost - ref.cost) / np.maximum(np.abs(ref.cost), 1.0)
            dx = np.abs(sol.x - ref.x).max(axis=1)
            ok = (relf <= COST_RTOL) & (dx <= CTRL_ATOL)
            same = (sol.nit == ref.nit) & (sol.nfev == ref.nfev) & (sol.status == ref.status)
            worst["ok"] = min(worst["ok"], ok.mean()); worst["same"] = min(worst["same"], same.mean())
            flag = "" if (ok.mean() >= 0.99 and ok[same].all() and same.mean() >= 0.85) else "  <-- FAIL"
            if flag:
                # is the oracle itself that sensitive here?  (one-ulp change of the start positions)
                pin = p0 if mode == "cold" else p0 + 0.1
                alt = oracle_mod.solve_batch(op, np.nextafter(pin, np.inf), v0, goal, nthreads=16,
                                             **({} if mode == "cold" else {"x_warm": np.nextafter(xw, np.inf)}))
                relf2 = np.abs(alt.cost - ref.cost) / np.maximum(np.abs(ref.cost), 1.0)
                ok2 = (relf2 <= COST_RTOL) & (np.abs(alt.x - ref.x).max(axis=1) <= CTRL_ATOL)
                same2 = (alt.nit == ref.nit) & (alt.nfev == ref.nfev) & (alt.status == ref.status)
                slack = 2.0 / np.sqrt(B)     # sampling noise of a fraction over B problems
                chaotic = ok.mean() >= ok2.mean() - 0.03 - slack and same.mean() >= same2.mean() - 0.06 - slack
                flag = (f"  <-- below the test thresholds; oracle vs itself after one ulp: ok={ok2.mean():.4f} "
                        f"same={same2.mean():.4f} -> {'chaotic configuration' if chaotic else 'FAIL'}")
                bad += 0 if chaotic else 1
            else:
                bad += 0
            print(f"seed {seed} trial {trial} {mode} N={N} B={B} maxit={kw['max_iterations']} ok={ok.mean():.4f} same={same.mean():.4f} "
                  f"maxdx_same={dx[same].max() if same.any() else 0:.2e} nit_max={ref.nit.max()}{flag}", flush=True)
print("worst", worst, "failures", bad)
