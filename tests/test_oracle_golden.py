"""The CPU oracle (oracle/*.c) against the golden fixtures produced by the UNMODIFIED
reference planner + SciPy 1.18.1 (tools/gen_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import SOLVER_FIXTURES, assert_solution_parity, golden_cfg, load_golden


@pytest.mark.parametrize("name", SOLVER_FIXTURES)
def test_solver_fixture(oracle_mod, name):
    d = load_golden(name)
    p = oracle_mod.make_params(horizon=int(d["N"]), dt=float(d["dt"]), **golden_cfg(d))
    xw = d["x_prev"] if "x_prev" in d.files else None
    r = oracle_mod.solve_batch(p, d["p0"], d["v0"], d["goal"], has_goal=d["has_goal"], x_warm=xw,
                               nthreads=4)
    assert_solution_parity(r, d, name)
    # the oracle is scalar fp64 like the reference: it should in fact agree to rounding
    assert np.nanmax(np.abs(r.x - d["x"])) < 1e-10


def test_named_values_of_survey_appendix_c(oracle_mod):
    """SURVEY.md App. C literal values (independent of the npz files)."""
    p = oracle_mod.make_params(horizon=6, dt=0.0025)
    r = oracle_mod.solve_batch(p, [[0, 0, 2]], [[0, 0, 0]], [[10, 0, 5]])
    assert (r.nit[0], r.nfev[0], r.status[0]) == (3, 5, 0)
    assert abs(r.cost[0] - 5233.477019206591) < 1e-8
    assert np.allclose(r.thrust_vectors[0][:, 2], 14.546524, atol=1e-6)
    assert np.allclose(r.attitudes[0][:, 2], -np.pi / 2)
    p = oracle_mod.make_params(horizon=8, dt=0.1)
    r = oracle_mod.solve_batch(p, [[0, 0, 2]], [[0, 0, 0]], [[10, 0, 5]])
    assert (r.nit[0], r.nfev[0]) == (3, 5) and abs(r.cost[0] - 4370.579984226455) < 1e-8
    # pinned defaults (reference tests/test_sitl_unit_tests.py:43-48)
    assert p.mass * p.gravity == pytest.approx(14.715)


def test_cold_start_and_clip(oracle_mod):
    for name in ("bench_N8_v", "default_N6", "no_goal_N6"):
        d = load_golden(name)
        p = oracle_mod.make_params(horizon=int(d["N"]), dt=float(d["dt"]))
        for b in range(0, len(d["p0"]), 17):
            goal = d["goal"][b] if d["has_goal"][b] else None
            x0 = oracle_mod.cold_start(p, d["p0"][b], d["v0"][b], goal)
            np.testing.assert_allclose(x0, d["x0"][b], rtol=0, atol=1e-12)


def test_warm_start_layout(oracle_mod):
    d = load_golden("warm_N8")
    p = oracle_mod.make_params(horizon=8, dt=0.1)
    for b in range(0, 96, 7):
        x0 = oracle_mod.warm_start(p, d["p0"][b], d["v0"][b], d["x_prev"][b])
        np.testing.assert_array_equal(x0, d["x0"][b])


def test_extract_so3(oracle_mod):
    d = load_golden("extract")
    p = oracle_mod.make_params(horizon=6, dt=float(d["G6_dt"]))
    x = np.concatenate([np.zeros(36), d["G6_T"].ravel()])
    acc, att, rates, thr = oracle_mod.extract(p, x)
    np.testing.assert_allclose(att, d["G6_att"], atol=1e-12)
    np.testing.assert_allclose(rates, d["G6_rates"], atol=1e-10)
    # literal values from SURVEY App. C G6
    assert att[1] == pytest.approx([-0.0712621853, -0.0356991127, -1.5707963268], abs=1e-9)
    assert rates[4] == pytest.approx([-0.9373514562, -2.3679804622, -0.5735875375], abs=1e-9)
    p = oracle_mod.make_params(horizon=8, dt=float(d["dt"]))
    for b in range(len(d["T"])):
        x = np.concatenate([np.zeros(48), d["T"][b].ravel()])
        acc, att, rates, thr = oracle_mod.extract(p, x)
        np.testing.assert_allclose(att, d["att"][b], atol=1e-12)
        np.testing.assert_allclose(rates, d["rates"][b], atol=1e-9)
        np.testing.assert_allclose(thr, d["thrusts"][b], atol=1e-12)
        np.testing.assert_allclose(acc, d["acc"][b], atol=1e-12)
