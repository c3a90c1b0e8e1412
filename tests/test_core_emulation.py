"""The kernel's own source (dart_planner_b200/csrc/se3mpc_core.cuh) compiled for the host with
a one-lane group, against the golden fixtures: exercises the exact control flow the CUDA
kernel runs (minus the shuffles) on the CPU-only tier.  Test infrastructure only."""
import numpy as np
import pytest

from conftest import SOLVER_FIXTURES, assert_solution_parity, golden_cfg, load_golden


def _params(d):
    from dart_planner_b200.config import SE3MPCConfig, make_params
    cfg = SE3MPCConfig(prediction_horizon=int(d["N"]), dt=float(d["dt"]), **golden_cfg(d))
    return make_params(cfg)


@pytest.mark.parametrize("name", SOLVER_FIXTURES)
def test_core_matches_reference_fixture(name):
    import emu
    d = load_golden(name)
    xw = d["x_prev"] if "x_prev" in d.files else None
    r = emu.solve_batch(_params(d), d["p0"], d["v0"], d["goal"], has_goal=d["has_goal"], x_warm=xw)
    assert_solution_parity(r, d, name)


def test_core_matches_oracle_random(oracle_mod):
    import emu
    from dart_planner_b200.config import SE3MPCConfig, make_params
    rng = np.random.default_rng(99)
    B = 3000
    p0 = rng.uniform(-10, 10, (B, 3))
    v0 = rng.uniform(-2, 2, (B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    goal[::7] = p0[::7] + rng.uniform(-0.05, 0.05, (len(p0[::7]), 3))   # near-goal population
    for N, dt in ((8, 0.1), (6, 0.0025)):
        ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=N, dt=dt), p0, v0, goal, nthreads=8)
        got = emu.solve_batch(make_params(SE3MPCConfig(prediction_horizon=N, dt=dt)), p0, v0, goal)
        same = (got.nit == ref.nit) & (got.nfev == ref.nfev) & (got.status == ref.status)
        assert same.mean() > 0.999, f"counter mismatches: {np.where(~same)[0][:10]}"
        dx = np.abs(got.x - ref.x).max(axis=1)
        assert (dx[same] < 1e-6).all()
