"""Row a13: `_compute_attitudes_and_rates` (se3_mpc_planner.py:604-654) on thrust sequences a
cold-started solve never produces -- zero-thrust steps (zero attitude, prev_R not advanced, so the
next valid step is differenced against the last VALID rotation over ONE dt), the degenerate
b1 = (1,0,0) x b3 fallback, tilted sequences -- against tests/golden/extract.npz, which
tools/gen_golden.py produced with the unmodified reference class (SURVEY G6 + 64 random cases).

CPU tier: the kernel core compiled for the host (tests/emu) and the host-side derivation of
solution rows (dart_planner_b200/derive.py).  GPU tier: the CUDA extraction through the C ABI
(`dart_se3mpc_extract_batch`) in every lane configuration, incl. two timesteps per lane."""
import os

import numpy as np
import pytest

from conftest import load_golden

ATT_TOL, RATE_TOL = 1e-12, 1e-9


def _cases():
    d = load_golden("extract")
    return d, float(d["dt"])


def _check(got, att, rates, thr, acc, what):
    a, at, r, th = got
    np.testing.assert_allclose(at, att, rtol=0, atol=ATT_TOL, err_msg=what)
    np.testing.assert_allclose(r, rates, rtol=0, atol=RATE_TOL, err_msg=what)
    np.testing.assert_allclose(th, thr, rtol=0, atol=1e-12, err_msg=what)
    np.testing.assert_allclose(a, acc, rtol=0, atol=1e-12, err_msg=what)


def _g6_expected(d):
    T = d["G6_T"]
    return d["G6_att"][None], d["G6_rates"][None], np.linalg.norm(T, axis=1)[None], (T / 1.5 - np.array([0, 0, 9.81]))[None]


def test_fixture_has_the_hard_cases():
    d, _ = _cases()
    mag = np.linalg.norm(d["T"], axis=2)
    assert (mag <= 1e-6).sum() >= 20                                     # zero-thrust steps
    assert (np.abs(d["T"] - np.array([3.0, 0, 0])).max(axis=2) == 0).sum() >= 10   # degenerate b1
    assert (mag[:, 0] <= 1e-6).any()                                     # an invalid FIRST step
    # G6: step 3 has zero thrust, step 4 is differenced against step 2 over one dt
    assert np.all(d["G6_att"][3] == 0) and np.all(d["G6_rates"][3] == 0) and np.any(d["G6_rates"][4] != 0)


def test_host_derivation_matches_the_reference():
    from dart_planner_b200.derive import derive_from_thrust
    d, dt = _cases()
    _check(derive_from_thrust(d["T"], dt, 1.5, 9.81), d["att"], d["rates"], d["thrusts"], d["acc"], "derive random")
    _check(derive_from_thrust(d["G6_T"][None], float(d["G6_dt"]), 1.5, 9.81), *_g6_expected(d), "derive G6")


def test_kernel_core_emulation_matches_the_reference():
    import emu
    import dart_planner_b200 as dp
    from dart_planner_b200.config import make_params
    d, dt = _cases()
    pr = make_params(dp.SE3MPCConfig(prediction_horizon=8, dt=dt))
    _check(emu.extract_batch(pr, d["T"]), d["att"], d["rates"], d["thrusts"], d["acc"], "emu random")
    p6 = make_params(dp.SE3MPCConfig(prediction_horizon=6, dt=float(d["G6_dt"])))
    _check(emu.extract_batch(p6, d["G6_T"][None]), *_g6_expected(d), "emu G6")
    # 7-slot instantiation: vertical thrust; closed form when every T_z > 1e-6, general path otherwise
    Tz = np.zeros((16, 8, 3))
    Tz[:, :, 2] = np.random.default_rng(0).uniform(2.0, 20.0, (16, 8))
    Tz[::3, 2, 2] = 0.0
    Tz[1, 0, 2] = 0.0
    _check(emu.extract_batch(pr, Tz, untilted=True), *emu.extract_batch(pr, Tz)[1:3], np.abs(Tz[:, :, 2]),
           Tz / 1.5 - np.array([0, 0, 9.81]), "emu untilted")


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [None, 1, 2, 3, 4])
def test_cuda_extraction_matches_the_reference(variant, monkeypatch):
    """Lane configurations: default (8 lanes at N = 8 / 6), 16, 32, 32 x 2 timesteps per lane."""
    import dart_planner_b200 as dp
    if variant is not None:
        monkeypatch.setenv("DART_SE3MPC_VARIANT", str(variant))
    d, dt = _cases()
    _check(dp.extract_batch(d["T"], dp.SE3MPCConfig(prediction_horizon=8, dt=dt)), d["att"], d["rates"], d["thrusts"],
           d["acc"], f"cuda random v{variant}")
    _check(dp.extract_batch(d["G6_T"][None], dp.SE3MPCConfig(prediction_horizon=6, dt=float(d["G6_dt"]))),
           *_g6_expected(d), f"cuda G6 v{variant}")


@pytest.mark.gpu
def test_cuda_extraction_four_lanes_and_long_horizons(oracle_mod, monkeypatch):
    import dart_planner_b200 as dp
    d, dt = _cases()
    # 4 lanes (N <= 4): the first four steps of the reference cases (later steps do not feed back)
    got = dp.extract_batch(d["T"][:, :4], dp.SE3MPCConfig(prediction_horizon=4, dt=dt))
    _check(got, d["att"][:, :4], d["rates"][:, :4], d["thrusts"][:, :4], d["acc"][:, :4], "cuda 4 lanes")
    # long horizons against the oracle (itself pinned by the fixture): 16, 32 lanes, 2 steps per lane
    rng = np.random.default_rng(9)
    for N in (13, 29, 40, 64):
        B = 96
        T = rng.normal(0, 4.0, (B, N, 3)) + np.array([0, 0, 12.0])
        T[rng.random((B, N)) < 0.15] = 0.0
        T[rng.random((B, N)) < 0.05] = np.array([3.0, 0, 0])
        T[0, :3] = 0.0
        T[1] = 0.0                                       # no valid step at all
        op = oracle_mod.make_params(horizon=N, dt=0.05)
        want = [oracle_mod.extract(op, np.concatenate([np.zeros(6 * N), T[b].ravel()])) for b in range(B)]
        acc, att, rates, thr = (np.array([w[i] for w in want]) for i in range(4))
        _check(dp.extract_batch(T, dp.SE3MPCConfig(prediction_horizon=N, dt=0.05)), att, rates, thr, acc, f"cuda N={N}")


@pytest.mark.gpu
@pytest.mark.parametrize("N", [5, 6])
def test_cuda_extraction_six_lane_groups(oracle_mod, monkeypatch, N):
    """6-lane groups (DART_SE3MPC_VARIANT=9, five problems per warp): invalid steps, the prev_R
    carry across lanes (group-relative ballot / broadcast) and the degenerate-b1 fallback against
    the oracle (pinned by the reference fixture), and the reference's G6 case at N = 6."""
    import dart_planner_b200 as dp
    monkeypatch.setenv("DART_SE3MPC_VARIANT", "9")
    rng = np.random.default_rng(19 + N)
    B = 103                                              # ragged: the last warp holds 3 of 5 groups
    T = rng.normal(0, 4.0, (B, N, 3)) + np.array([0, 0, 12.0])
    T[rng.random((B, N)) < 0.2] = 0.0
    T[rng.random((B, N)) < 0.08] = np.array([3.0, 0, 0])
    T[0, :3] = 0.0
    T[1] = 0.0
    op = oracle_mod.make_params(horizon=N, dt=0.05)
    want = [oracle_mod.extract(op, np.concatenate([np.zeros(6 * N), T[b].ravel()])) for b in range(B)]
    acc, att, rates, thr = (np.array([w[i] for w in want]) for i in range(4))
    _check(dp.extract_batch(T, dp.SE3MPCConfig(prediction_horizon=N, dt=0.05)), att, rates, thr, acc, f"cuda 6 lanes N={N}")
    if N == 6:
        d, _ = _cases()
        _check(dp.extract_batch(d["G6_T"][None], dp.SE3MPCConfig(prediction_horizon=6, dt=float(d["G6_dt"]))),
               *_g6_expected(d), "cuda G6, 6 lanes")


@pytest.mark.gpu
@pytest.mark.parametrize("N", [8, 6, 13, 40])
def test_cuda_solve_epilogue_with_invalid_steps(N, oracle_mod):
    """The same branches through the public solve: min_thrust = 0 and a warm start whose shifted
    last thrust is 0 (it stays 0 in the reference gradient mode when w_thrust pulls nothing), plus
    zero thrusts placed mid-horizon -- derived rows of the solve against the oracle's extraction of
    the kernel's own x."""
    import dart_planner_b200 as dp
    rng = np.random.default_rng(40 + N)
    B = 128
    p0 = rng.uniform(-5, 5, (B, 3)); v0 = rng.uniform(-1, 1, (B, 3))
    goal = p0 + rng.uniform(-3, 3, (B, 3))
    cfg = dp.SE3MPCConfig(prediction_horizon=N, dt=0.1, min_thrust=0.0, thrust_weight=0.0, max_iterations=1)
    xw = np.zeros((B, 9 * N))
    xw[:, : 3 * N] = np.repeat(p0, N, axis=0).reshape(B, 3 * N)
    T = rng.normal(0, 2.0, (B, N, 3)) + np.array([0, 0, 14.0])
    T[rng.random((B, N)) < 0.2] = 0.0
    xw[:, 6 * N:] = T.reshape(B, 3 * N)
    got = dp.plan_batch(p0, v0, goal, cfg, x_warm=xw, to_host=True)
    thr = np.linalg.norm(got.thrust_vectors, axis=2)
    assert (thr <= 1e-6).sum() >= B              # at least the shifted-in last step of every problem
    op = oracle_mod.make_params(horizon=N, dt=0.1, min_thrust=0.0, thrust_weight=0.0, max_iterations=1)
    want = [oracle_mod.extract(op, got.x[b]) for b in range(B)]
    acc, att, rates, th = (np.array([w[i] for w in want]) for i in range(4))
    _check((got.accelerations, got.attitudes, got.body_rates, got.thrusts), att, rates, th, acc, f"solve N={N}")
