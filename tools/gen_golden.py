#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference planner and mapper.

Runs only in the build container (needs /root/reference).  The reference is imported
as-is from /root/reference/src with the unit-transparent `pint` stand-in of
tools/refshim (pint is not installed), and drives the installed SciPy
(L-BFGS-B).  `scipy.optimize.minimize` is observed (not altered) through a recording
wrapper bound in the planner module's namespace so that fun / nit / nfev / status are
captured next to the planner's own outputs.

    python tools/gen_golden.py            # rewrites tests/golden/

The fixtures are small and committed; nothing on the GPU box reads /root/reference.
"""
from __future__ import annotations

import logging
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DART_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, os.path.join(REF, "src"))
logging.disable(logging.CRITICAL)

import scipy  # noqa: E402
from dart_planner.common import timing_alignment as ta  # noqa: E402
from dart_planner.common.types import DroneState  # noqa: E402
from dart_planner.perception.explicit_geometric_mapper import ExplicitGeometricMapper  # noqa: E402
from dart_planner.planning import se3_mpc_planner as ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
_last = {}
_orig_minimize = ref.minimize


def _recording_minimize(*a, **k):
    res = _orig_minimize(*a, **k)
    _last["res"] = res
    _last["x0"] = np.array(k.get("x0", a[1] if len(a) > 1 else None), dtype=float)
    return res


ref.minimize = _recording_minimize


def make_planner(N, dt, **cfg):
    """Planner built through the reference's own mechanisms."""
    ta.reset_timing_manager()
    ta.get_timing_manager(ta.TimingConfig(control_frequency=1.0 / dt))
    pl = ref.SE3MPCPlanner(ref.SE3MPCConfig(prediction_horizon=N, **cfg))
    assert abs(pl.se3_config.dt - dt) < 1e-15
    return pl


def solve_cases(N, dt, p0, v0, goal, has_goal=None, x_prev=None, **cfg):
    """Run the reference on each row; returns dict of stacked arrays."""
    pl = make_planner(N, dt, **cfg)
    B = len(p0)
    n = 9 * N
    out = dict(
        N=np.int32(N), dt=np.float64(dt), p0=np.asarray(p0, float), v0=np.asarray(v0, float),
        goal=np.asarray(goal, float),
        has_goal=np.ones(B, np.uint8) if has_goal is None else np.asarray(has_goal, np.uint8),
        x=np.zeros((B, n)), x0=np.zeros((B, n)), fun=np.zeros(B), nit=np.zeros(B, np.int32),
        nfev=np.zeros(B, np.int32), status=np.zeros(B, np.int32), success=np.zeros(B, np.uint8),
        accelerations=np.zeros((B, N, 3)), attitudes=np.zeros((B, N, 3)),
        body_rates=np.zeros((B, N, 3)), thrusts=np.zeros((B, N)),
    )
    if x_prev is not None:
        out["x_prev"] = np.asarray(x_prev, float)
    for k, v in cfg.items():
        out["cfg_" + k] = np.float64(v)
    for b in range(B):
        st = DroneState(timestamp=0.0, position=np.array(p0[b], float),
                        velocity=np.array(v0[b], float))
        pl.goal_position = None          # defeat the 0.5 m hysteresis (:199)
        pl.last_solution = None
        if out["has_goal"][b]:
            pl.sense(st, np.array(goal[b], float))
        if x_prev is not None:
            P, V, T = pl._unpack_variables(np.array(x_prev[b], float), N)
            pl.last_solution = {"positions": P.copy(), "velocities": V.copy(),
                                "thrust_vectors": T.copy()}
        sol = pl.plan(st)
        r = _last["res"]
        out["x"][b] = r.x
        out["x0"][b] = _last["x0"]
        out["fun"][b] = r.fun
        out["nit"][b] = r.nit
        out["nfev"][b] = r.nfev
        out["status"][b] = r.status
        out["success"][b] = bool(r.success)
        out["accelerations"][b] = sol["accelerations"]
        out["attitudes"][b] = sol["attitudes"]
        out["body_rates"][b] = sol["body_rates"]
        out["thrusts"][b] = sol["thrusts"]
        assert np.array_equal(np.asarray(sol["thrust_vectors"]).ravel(), r.x[6 * N:], equal_nan=True)
    return out


def bench_inputs(rng, B, v_scale=0.0):
    """SURVEY 8d config 2/3 distribution (experiments/validation/benchmark_audit_improvements.py:292-302)."""
    p0 = rng.uniform(-10, 10, (B, 3))
    v0 = rng.uniform(-v_scale, v_scale, (B, 3)) if v_scale > 0 else np.zeros((B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    return p0, v0, goal


def save(name, d):
    os.makedirs(OUT, exist_ok=True)
    d = dict(d)
    d["scipy_version"] = np.array(scipy.__version__)
    d["numpy_version"] = np.array(np.__version__)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"wrote {name}.npz  ({', '.join(k for k in d if not k.endswith('_version'))})")


def gen_solver():
    # --- named cases of SURVEY App. C ---------------------------------------------------
    save("G1", solve_cases(6, 0.0025, [[0, 0, 2]], [[0, 0, 0]], [[10, 0, 5]]))
    save("G2", solve_cases(8, 0.1, [[0, 0, 2]], [[0, 0, 0]], [[10, 0, 5]]))
    save("G3", solve_cases(8, 0.1, [[1, -2, 3]], [[0.5, 0.2, -0.1]], [[4, 1, 5]]))
    save("G4", solve_cases(6, 0.0025, [[1, -2, 3]], [[0, 0, 0]], [[1, -2, 3]]))
    save("G5", solve_cases(6, 0.0025, [[1, -2, 3]], [[0.5, 0.2, -0.1]], [[1.004, -2.003, 3.002]]))

    # --- random regimes -----------------------------------------------------------------
    rng = np.random.default_rng(1)
    save("bench_N8", solve_cases(8, 0.1, *bench_inputs(rng, 256)))
    rng = np.random.default_rng(2)
    save("bench_N8_v", solve_cases(8, 0.1, *bench_inputs(rng, 256, 2.0)))
    rng = np.random.default_rng(11)
    save("default_N6", solve_cases(6, 0.0025, *bench_inputs(rng, 256, 2.0)))
    rng = np.random.default_rng(12)
    p0 = rng.uniform(-5, 5, (96, 3))
    save("near_goal_N6", solve_cases(6, 0.0025, p0, np.zeros((96, 3)),
                                     p0 + rng.uniform(-0.01, 0.01, (96, 3))))
    rng = np.random.default_rng(13)
    p0 = rng.uniform(-5, 5, (64, 3))
    save("near_goal_N8_v", solve_cases(8, 0.1, p0, rng.uniform(-0.5, 0.5, (64, 3)),
                                       p0 + rng.uniform(-0.3, 0.3, (64, 3))))
    rng = np.random.default_rng(14)
    p0 = rng.uniform(-5, 5, (8, 3))
    save("at_goal_N8", solve_cases(8, 0.1, p0, np.zeros((8, 3)), p0.copy()))
    # Monte-Carlo config 4 distribution
    rng = np.random.default_rng(3)
    p0 = np.array([0, 0, 2.0]) + rng.normal(0, 1.0, (128, 3))
    save("monte_carlo_N8", solve_cases(8, 0.1, p0, rng.normal(0, 0.5, (128, 3)),
                                       np.tile([10.0, 0, 5.0], (128, 1))))
    # no goal (plan() with goal_position None, :352-355, :523, :566)
    rng = np.random.default_rng(15)
    save("no_goal_N6", solve_cases(6, 0.0025, rng.uniform(-5, 5, (16, 3)),
                                   rng.uniform(-3, 3, (16, 3)), np.zeros((16, 3)),
                                   has_goal=np.zeros(16, np.uint8)))
    # other horizons
    rng = np.random.default_rng(16)
    save("bench_N20", solve_cases(20, 0.1, *bench_inputs(rng, 32, 2.0)))
    rng = np.random.default_rng(17)
    save("bench_N4", solve_cases(4, 0.05, *bench_inputs(rng, 32, 2.0)))
    rng = np.random.default_rng(23)
    save("bench_N13", solve_cases(13, 0.1, *bench_inputs(rng, 32, 2.0)))
    # config knobs (SURVEY App. B sensitivity table)
    rng = np.random.default_rng(18)
    save("tol0p1_it5_N6", solve_cases(6, 0.0025, *bench_inputs(rng, 48, 2.0),
                                      convergence_tolerance=0.1, max_iterations=5))
    rng = np.random.default_rng(19)
    save("tol1em4_N6", solve_cases(6, 0.0025, *bench_inputs(rng, 96, 2.0),
                                   convergence_tolerance=1e-4, max_iterations=15))
    rng = np.random.default_rng(20)
    save("tol1em8_it40_N8", solve_cases(8, 0.1, *bench_inputs(rng, 48, 2.0),
                                        convergence_tolerance=1e-8, max_iterations=40))
    rng = np.random.default_rng(21)
    save("maxiter2_N8", solve_cases(8, 0.1, *bench_inputs(rng, 48, 2.0),
                                    convergence_tolerance=1e-3, max_iterations=2))
    rng = np.random.default_rng(24)
    save("weights_N8", solve_cases(8, 0.1, *bench_inputs(rng, 48, 2.0), position_weight=37.0,
                                   velocity_weight=3.5, thrust_weight=0.7,
                                   acceleration_weight=2.0, max_velocity=6.0,
                                   max_thrust=20.0, min_thrust=1.0, max_tilt_angle=0.6))

    # --- warm start (:294-327; dead in the reference because last_solution is never
    # assigned, exercised here by assigning it the way the code reads it) -------------
    rng = np.random.default_rng(22)
    p0, v0, goal = bench_inputs(rng, 96, 2.0)
    first = solve_cases(8, 0.1, p0, v0, goal)
    m, g = 1.5, 9.81
    T0 = first["x"][:, 48:51]
    a0 = T0 / m - np.array([0, 0, g])
    p1 = p0 + v0 * 0.1 + 0.5 * a0 * 0.01
    v1 = v0 + a0 * 0.1
    # tilt some of the previous thrusts so attitudes/body rates are non-trivial
    xprev = first["x"].copy()
    xprev[::3, 48:] += rng.normal(0, 1.0, xprev[::3, 48:].shape)
    save("warm_N8", solve_cases(8, 0.1, p1, v1, goal, x_prev=xprev))


def gen_nonfinite():
    """Non-finite states / goals / previous solutions: a NaN survives SciPy's clip of x0,
    every line-search trial is rejected and the routine ends ABNORMAL at the start point (nit 0,
    status 2, fun NaN, nfev = maxls + 1, + 1 when x itself holds a NaN: SciPy's evaluation cache
    compares x by value); infinities are clipped to the bounds where the guess keeps them
    infinite and become NaN where the guess forms inf - inf or 0 * inf."""
    nan, inf = float("nan"), float("inf")
    base = ([0.0, 0.0, 2.0], [0.0, 0.0, 0.0], [10.0, 0.0, 5.0])
    cases = []
    for which in range(3):
        for comp in range(3):
            for val in (nan, inf, -inf):
                c = [list(base[0]), list(base[1]), list(base[2])]
                c[which][comp] = val
                cases.append(c)
    p0 = np.array([c[0] for c in cases]); v0 = np.array([c[1] for c in cases]); goal = np.array([c[2] for c in cases])
    with np.errstate(all="ignore"):
        save("nonfinite_N8", solve_cases(8, 0.1, p0, v0, goal))
        save("nonfinite_N6", solve_cases(6, 0.0025, p0[::2], v0[::2], goal[::2]))
        # warm starts: NaN / inf inside the previous solution, NaN goal with a finite guess
        first = solve_cases(8, 0.1, [base[0]], [base[1]], [base[2]])
        xp = np.repeat(first["x"], 8, axis=0)
        xp[1, 60] = nan          # T_4.x  (shifted into T_3)
        xp[2, 5] = nan           # P_1.z
        xp[3, 30] = inf          # V_2.x  (clipped to the bound)
        xp[4, 71] = nan          # T_7.z
        xp[5, 48] = nan          # T_0.x is dropped by the shift: a clean solve
        xp[6, 0] = nan           # P_0 is replaced by the state: a clean solve
        gw = np.tile(base[2], (8, 1)); gw[7, 1] = nan
        save("nonfinite_warm_N8", solve_cases(8, 0.1, np.tile(base[0], (8, 1)), np.tile(base[1], (8, 1)), gw, x_prev=xp))


def gen_extract():
    pl = make_planner(6, 0.1)
    T = np.array([(0, 0, 14.715), (1, 0.5, 14), (2, -1, 13), (0, 0, 0), (3, 2, 12),
                  (-1.5, 0.75, 15)], float)
    att, rates = pl._compute_attitudes_and_rates(T, np.zeros((6, 3)))
    d = dict(G6_T=T, G6_att=att, G6_rates=rates, G6_dt=np.float64(0.1))
    rng = np.random.default_rng(5)
    B, N = 64, 8
    pl = make_planner(N, 0.1)
    Ts = rng.normal(0, 4.0, (B, N, 3)) + np.array([0, 0, 12.0])
    Ts[rng.random((B, N)) < 0.1] = 0.0                       # low-thrust steps (prev_R skip)
    Ts[rng.random((B, N)) < 0.05] = np.array([3.0, 0, 0])    # b1 degenerate (yaw x b3 = 0)
    Ts[0, 0] = 0.0
    atts, ratess, thr, acc = [], [], [], []
    for b in range(B):
        x = np.concatenate([np.zeros(6 * N), Ts[b].ravel()])
        sol = pl._extract_solution_from_result(x, N)
        atts.append(sol["attitudes"]); ratess.append(sol["body_rates"])
        thr.append(sol["thrusts"]); acc.append(sol["accelerations"])
    d.update(T=Ts, att=np.array(atts), rates=np.array(ratess), thrusts=np.array(thr),
             acc=np.array(acc), dt=np.float64(0.1))
    save("extract", d)


def gen_mapper():
    d = {}
    mp = ExplicitGeometricMapper(resolution=0.5)
    k1 = mp._trace_ray(np.zeros(3), np.array([1.0, 1.0, 0.0]) / np.sqrt(2), 5.0)
    k2 = mp._trace_ray(np.array([0.3, -0.2, 1.1]), np.array([-1.0, 2.0, 0.5]), 3.0)
    d["kat1"] = np.array(k1, np.int32)
    d["kat2"] = np.array(k2, np.int32)
    # random rays at the config-3 resolution
    rng = np.random.default_rng(7)
    mp2 = ExplicitGeometricMapper(resolution=0.2)
    R = 200
    starts = rng.uniform(-20, 20, (R, 3))
    dirs = rng.normal(0, 1, (R, 3))
    dirs[:10, 2] = 0.0
    dirs[10:20, 1:] = 0.0
    dists = rng.uniform(0.05, 20, R)
    lens, vox = [], []
    for i in range(R):
        v = np.array(mp2._trace_ray(starts[i], dirs[i], dists[i]), np.int32).reshape(-1, 3)
        lens.append(len(v)); vox.append(v)
    d.update(ray_res=np.float64(0.2), ray_start=starts, ray_dir=dirs, ray_dist=dists,
             ray_len=np.array(lens, np.int32), ray_vox=np.concatenate(vox, 0))
    # KAT3 sphere + safety
    mp3 = ExplicitGeometricMapper(resolution=0.2)
    mp3.add_obstacle(np.array([15.0, 5.0, 5.0]), 2.0)
    d["kat3_nvox"] = np.int32(len(mp3.voxels))
    q = np.array([(15, 5, 5), (16.9, 5, 5), (17.15, 5, 5), (0, 0, 0)], float)
    d["kat3_q"] = q
    d["kat3_occ"] = np.array([mp3.query_occupancy(x) for x in q])
    traj = np.linspace((5, 5, 5), (20, 5, 5), 8)
    safe, idx = mp3.is_trajectory_safe(traj, 1.5, 0.6)
    d["kat3_traj"] = traj
    d["kat3_safe"] = np.array([int(bool(safe)), idx], np.int32)
    # random spheres + random queries + trajectory checks
    rng = np.random.default_rng(8)
    mp4 = ExplicitGeometricMapper(resolution=0.2)
    cs = rng.uniform(-8, 8, (12, 3)); rs = rng.uniform(0.5, 2.0, 12)
    for c, r in zip(cs, rs):
        mp4.add_obstacle(c, float(r))
    qs = rng.uniform(-11, 11, (2000, 3))
    d.update(sph_c=cs, sph_r=rs, sph_q=qs, sph_occ=np.array([mp4.query_occupancy(x) for x in qs]),
             sph_nvox=np.int32(len(mp4.voxels)))
    trajs = rng.uniform(-10, 10, (200, 1, 3)) + np.cumsum(rng.normal(0, 0.6, (200, 8, 3)), axis=1)
    res = [mp4.is_trajectory_safe(t, 1.5, 0.6) for t in trajs]
    d.update(sph_traj=trajs, sph_safe=np.array([int(bool(s)) for s, _ in res], np.int32),
             sph_idx=np.array([i for _, i in res], np.int32))
    # Bayes sequence (KAT4)
    mp5 = ExplicitGeometricMapper(resolution=0.2)
    from dart_planner.perception.explicit_geometric_mapper import VoxelData
    vx, seq = VoxelData(), []
    assert vx.occupancy_probability == 0.5
    for h in [1, 1, 0, 1, 0, 0, 0]:
        mp5._bayesian_update(vx, bool(h)); seq.append(float(vx.occupancy_probability))
    d["kat4"] = np.array(seq)
    save("mapper", d)


def gen_update_map():
    """update_map (explicit_geometric_mapper.py:100-152) on LiDAR-like scans: the resulting
    sparse dict of occupancy probabilities (keys + values) and the returned counters."""
    from dart_planner.perception.explicit_geometric_mapper import SensorObservation
    rng = np.random.default_rng(11)
    mp = ExplicitGeometricMapper(resolution=0.5, max_range=12.0)
    R = 600
    sensors = np.array([[0.2, 0.3, 1.0], [4.1, -2.2, 2.5], [-3.3, 3.7, 1.4]])
    pos = sensors[rng.integers(0, 3, R)]
    dirs = rng.normal(0, 1, (R, 3))
    dirs[:30, 2] = 0.0                               # planar sweep
    hit = rng.uniform(0.3, 15.0, R)                  # some beyond the mapper's max_range
    hit[rng.random(R) < 0.3] = np.nan                # no return -> max_range, endpoint is a miss
    hit[5] = 0.0                                     # falsy hit distance quirk (:112)
    mr = np.full(R, 10.0)
    mr[::7] = 20.0
    obs = [SensorObservation(position=pos[i], direction=dirs[i],
                             hit_distance=None if np.isnan(hit[i]) else float(hit[i]),
                             max_range=float(mr[i]), timestamp=float(i)) for i in range(R)]
    out = mp.update_map(obs[: R // 2])
    out2 = mp.update_map(obs[R // 2:])               # second batch on top of the first
    keys = np.array(sorted(mp.voxels), np.int32)
    probs = np.array([mp.voxels[tuple(k)].occupancy_probability for k in keys])
    counts = np.array([mp.voxels[tuple(k)].observation_count for k in keys], np.int32)
    save("update_map", dict(res=np.float64(0.5), mapper_max_range=np.float64(12.0), pos=pos, dir=dirs,
                            hit=hit, obs_max_range=mr, split=np.int32(R // 2), keys=keys, probs=probs,
                            counts=counts, updated=np.array([out["updated_voxels"], out2["updated_voxels"]],
                                                            np.int64)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "update_map":
        gen_update_map()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "nonfinite":
        gen_nonfinite()
        sys.exit(0)
    gen_solver()
    gen_nonfinite()
    gen_extract()
    gen_mapper()
    gen_update_map()
