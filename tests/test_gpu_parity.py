"""GPU parity tests proper: the CUDA path, called through the C ABI, against
(1) the golden fixtures produced by the unmodified reference, (2) the oracle on seeded random
batches, (3) size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

from conftest import (COST_RTOL, CTRL_ATOL, SOLVER_FIXTURES, assert_solution_parity, golden_cfg,
                      load_golden)

pytestmark = pytest.mark.gpu


def _cfg(d):
    from dart_planner_b200 import SE3MPCConfig
    return SE3MPCConfig(prediction_horizon=int(d["N"]), dt=float(d["dt"]), **golden_cfg(d))


@pytest.mark.parametrize("name", SOLVER_FIXTURES)
def test_gpu_matches_reference_fixture(name):
    import dart_planner_b200 as dp
    d = load_golden(name)
    xw = d["x_prev"] if "x_prev" in d.files else None
    sol = dp.plan_batch(d["p0"], d["v0"], d["goal"], _cfg(d), has_goal=d["has_goal"], x_warm=xw,
                        to_host=True)
    assert_solution_parity(sol, d, name)


def bench_inputs(seed, B, v_scale=0.0):
    rng = np.random.default_rng(seed)
    p0 = rng.uniform(-10, 10, (B, 3))
    v0 = rng.uniform(-v_scale, v_scale, (B, 3)) if v_scale > 0 else np.zeros((B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    return p0, v0, goal


def _compare(sol, ref, min_counter_agreement=1.0):
    relf = np.abs(sol.cost - ref.cost) / np.maximum(np.abs(ref.cost), 1.0)
    dx = np.abs(sol.x - ref.x).max(axis=1)
    same = (sol.nit == ref.nit) & (sol.nfev == ref.nfev) & (sol.status == ref.status)
    assert same.mean() >= min_counter_agreement, \
        f"counters differ on {(~same).sum()} of {same.size}: first {np.where(~same)[0][:8]}"
    ok = (relf <= COST_RTOL) & (dx <= CTRL_ATOL)
    assert ok[same].all(), f"{(~ok[same]).sum()} problems out of tolerance with equal counters"
    assert ok.mean() >= min_counter_agreement
    np.testing.assert_allclose(sol.attitudes[same], ref.attitudes[same], atol=1e-6)
    np.testing.assert_allclose(sol.thrusts[same], ref.thrusts[same], atol=1e-4)
    np.testing.assert_allclose(sol.accelerations[same], ref.accelerations[same], atol=1e-4)


@pytest.mark.parametrize("N,dt,seed,vs", [(8, 0.1, 1, 0.0), (8, 0.1, 2, 2.0), (6, 0.0025, 3, 2.0),
                                          (4, 0.05, 4, 1.0), (13, 0.1, 5, 2.0), (20, 0.1, 6, 2.0),
                                          (32, 0.1, 7, 2.0), (40, 0.1, 8, 2.0)])
def test_gpu_matches_oracle_random(oracle_mod, N, dt, seed, vs):
    """BASELINE configs[1] inputs (seed 1: 4096 hover-to-goal solves) and other horizons."""
    import dart_planner_b200 as dp
    B = 4096 if N <= 8 else 1024
    p0, v0, goal = bench_inputs(seed, B, vs)
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=N, dt=dt), p0, v0, goal, nthreads=16)
    sol = dp.plan_batch(p0, v0, goal, dp.SE3MPCConfig(prediction_horizon=N, dt=dt), to_host=True)
    _compare(sol, ref)


@pytest.mark.parametrize("N,seed", [(13, 41), (32, 42)])
def test_gpu_throughput_builds_wide_lanes(oracle_mod, N, seed):
    """>= 12288 problems at 8 < N <= 32 run on the 168-register builds of the 16- / 32-lane
    configurations."""
    import ctypes as C
    import dart_planner_b200 as dp
    from dart_planner_b200 import _cabi
    from dart_planner_b200.config import make_params
    B = 12288 + 40
    regs = C.c_int32()
    _cabi.lib().dart_se3mpc_kernel_info(C.byref(make_params(dp.SE3MPCConfig(prediction_horizon=N, dt=0.1))), B,
                                        None, None, None, None, C.byref(regs))
    assert regs.value <= 168
    p0, v0, goal = bench_inputs(seed, B, 2.0)
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=N, dt=0.1), p0, v0, goal, nthreads=16)
    sol = dp.plan_batch(p0, v0, goal, dp.SE3MPCConfig(prediction_horizon=N, dt=0.1), to_host=True)
    _compare(sol, ref)
    xw = ref.x.copy()
    xw[:, 6 * N:] += np.random.default_rng(seed).normal(0, 0.4, xw[:, 6 * N:].shape)
    ref_w = oracle_mod.solve_batch(oracle_mod.make_params(horizon=N, dt=0.1), p0, v0, goal, x_warm=xw, nthreads=16)
    sol_w = dp.plan_batch(p0, v0, goal, dp.SE3MPCConfig(prediction_horizon=N, dt=0.1), x_warm=xw, to_host=True)
    _compare(sol_w, ref_w, min_counter_agreement=0.995)


def test_gpu_near_goal_regime(oracle_mod):
    """Degenerate regime (line search fails, ABNORMAL endings): decisions sit on rounding
    noise, so a small fraction of counter differences is tolerated and reported."""
    import dart_planner_b200 as dp
    rng = np.random.default_rng(12)
    B = 2048
    p0 = rng.uniform(-5, 5, (B, 3))
    goal = p0 + rng.uniform(-0.01, 0.01, (B, 3))
    goal[:64] = p0[:64]
    v0 = np.zeros((B, 3))
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=6, dt=0.0025), p0, v0, goal, nthreads=16)
    sol = dp.plan_batch(p0, v0, goal, dp.SE3MPCConfig(prediction_horizon=6, dt=0.0025), to_host=True)
    _compare(sol, ref, min_counter_agreement=0.97)


def _stationary_warm_starts(rng, N, B, min_thrust=0.5):
    """Warm starts that sit (almost) on the stationary point of the reference's gradient: P = goal,
    V = 0, T = (0, 0, min_thrust) -- the projected gradient is ~0, the solve stops at [401]."""
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    xw = np.zeros((B, 9 * N))
    xw[:, :3 * N] = np.tile(goal, (1, N)) + rng.normal(0, 1e-5, (B, 3 * N))
    xw[:, 6 * N + 2::3] = min_thrust
    return goal, xw


@pytest.mark.parametrize("N", [3, 5, 6, 8, 13, 20])
def test_gpu_projected_gradient_stop_with_unused_lanes(oracle_mod, N):
    """Horizons that leave lanes (or timestep slots) of the group unused: the phantom T_z slots hold
    0, outside [min_thrust, max_thrust], and must not enter the projected-gradient norm (round-1
    kernel: they did, so a [401] stop was never taken at N = 5, 6, 7, ...)."""
    import dart_planner_b200 as dp
    rng = np.random.default_rng(60 + N)
    goal, xw = _stationary_warm_starts(rng, N, 300)
    v0 = np.zeros_like(goal)
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=N, dt=0.1), goal, v0, goal, x_warm=xw, nthreads=16)
    assert (ref.task == 1).all() and (ref.nit == 0).all()       # CONV_PGTOL at the start
    sol = dp.plan_batch(goal, v0, goal, dp.SE3MPCConfig(prediction_horizon=N, dt=0.1), x_warm=xw, to_host=True)
    _compare(sol, ref)
    assert (sol.nit == 0).all()


def test_gpu_warm_start_and_masks(oracle_mod):
    import dart_planner_b200 as dp
    p0, v0, goal = bench_inputs(22, 2048, 2.0)
    cfg = dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)
    first = dp.plan_batch(p0, v0, goal, cfg, to_host=True)
    rng = np.random.default_rng(5)
    xprev = first.x.copy()
    xprev[::3, 48:] += rng.normal(0, 1.0, xprev[::3, 48:].shape)   # tilt some thrusts
    a0 = first.x[:, 48:51] / 1.5 - np.array([0, 0, 9.81])
    p1, v1 = p0 + 0.1 * v0 + 0.005 * a0, v0 + 0.1 * a0
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=8, dt=0.1), p1, v1, goal, x_warm=xprev,
                                 nthreads=16)
    sol = dp.plan_batch(p1, v1, goal, cfg, x_warm=xprev, to_host=True)
    _compare(sol, ref)
    assert np.abs(sol.body_rates).max() > 0.1       # the SO(3) path is exercised
    np.testing.assert_allclose(sol.body_rates, ref.body_rates, atol=1e-5, rtol=1e-6)
    # warm_mask: even problems warm, odd problems cold
    mask = (np.arange(len(p1)) % 2 == 0).astype(np.uint8)
    cold = oracle_mod.solve_batch(oracle_mod.make_params(horizon=8, dt=0.1), p1, v1, goal, nthreads=16)
    mix = dp.plan_batch(p1, v1, goal, cfg, x_warm=xprev, warm_mask=mask, to_host=True)
    np.testing.assert_allclose(mix.x[0::2], ref.x[0::2], atol=1e-6)
    np.testing.assert_allclose(mix.x[1::2], cold.x[1::2], atol=1e-6)
    # has_goal mask
    hg = (np.arange(len(p1)) % 3 != 0).astype(np.uint8)
    ref2 = oracle_mod.solve_batch(oracle_mod.make_params(horizon=8, dt=0.1), p1, v1, goal, has_goal=hg,
                                  nthreads=16)
    got2 = dp.plan_batch(p1, v1, goal, cfg, has_goal=hg, to_host=True)
    _compare(got2, ref2, min_counter_agreement=0.98)


def test_gpu_consistent_gradient_extension(oracle_mod):
    """gradient_mode=1 (exact gradient; extension, self-oracle): many iterations, the
    correction ring fills and the maxiter stop is reached."""
    import dart_planner_b200 as dp
    p0, v0, goal = bench_inputs(31, 1024, 2.0)
    for maxiter, tol in ((15, 1e-6), (40, 1e-9)):
        ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=8, dt=0.1, consistent_gradient=1,
                                                            max_iterations=maxiter,
                                                            convergence_tolerance=tol), p0, v0, goal,
                                     nthreads=16)
        sol = dp.plan_batch(p0, v0, goal, dp.SE3MPCConfig(prediction_horizon=8, dt=0.1,
                                                          max_iterations=maxiter,
                                                          convergence_tolerance=tol),
                            gradient_mode=1, to_host=True)
        assert ref.nit.max() >= 5
        _compare(sol, ref, min_counter_agreement=0.98)


def test_gpu_edge_sizes():
    import dart_planner_b200 as dp
    cfg = dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)
    for B in (1, 3, 5, 31, 33, 127, 4097):
        p0, v0, goal = bench_inputs(B, B, 1.0)
        sol = dp.plan_batch(p0, v0, goal, cfg, to_host=True)
        one = dp.plan_batch(p0[-1:], v0[-1:], goal[-1:], cfg, to_host=True)
        np.testing.assert_array_equal(sol.x[-1], one.x[0])          # independent of batch position
        assert sol.nit[-1] == one.nit[0]


def test_full_size_properties():
    """65 536 problems (BASELINE configs[2] inputs, seed 2): properties that need no oracle."""
    import dart_planner_b200 as dp
    B = 65536
    p0, v0, goal = bench_inputs(2, B, 2.0)
    cfg = dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)
    sol = dp.plan_batch(p0, v0, goal, cfg, to_host=True)
    lo = np.r_[np.full(24, -100.0), np.full(24, -10.0), np.tile([-17.67766952966369, -17.67766952966369, 2.0], 8)]
    hi = np.r_[np.full(24, 100.0), np.full(24, 10.0), np.tile([17.67766952966369, 17.67766952966369, 25.0], 8)]
    assert (sol.x >= lo - 1e-12).all() and (sol.x <= hi + 1e-12).all()     # feasibility
    assert np.isfinite(sol.x).all() and np.isfinite(sol.cost).all()
    assert (sol.nfev >= sol.nit + 1).all() and (sol.nit <= 15).all() and np.isin(sol.status, (0, 1, 2)).all()
    # determinism + permutation equivariance (problems are independent)
    perm = np.random.default_rng(0).permutation(B)
    sol2 = dp.plan_batch(p0[perm], v0[perm], goal[perm], cfg, to_host=True)
    np.testing.assert_array_equal(sol2.x, sol.x[perm])
    np.testing.assert_array_equal(sol2.nfev, sol.nfev[perm])
    # translation equivariance in xy... does not hold (bounds are absolute); instead: the
    # reported cost equals the reference objective (:516-550) evaluated at the reported x
    P, V, T = sol.positions, sol.velocities, sol.thrust_vectors
    e = P - goal[:, None, :]
    f = (100 * (e ** 2).sum((1, 2)) + 10 * (V ** 2).sum((1, 2))
         + ((T / 1.5 - np.array([0, 0, 9.81])) ** 2).sum((1, 2))
         + 0.1 * ((T - np.array([0, 0, 14.715])) ** 2).sum((1, 2)) + 1000 * (e[:, -1] ** 2).sum(1))
    ok = sol.status != 2          # ABNORMAL returns the restored iterate but the last evaluated f
    np.testing.assert_allclose(sol.cost[ok], f[ok], rtol=1e-10)
    # derived outputs are consistent with x
    np.testing.assert_allclose(sol.thrusts, np.linalg.norm(T, axis=2), rtol=1e-12)
    np.testing.assert_allclose(sol.accelerations, T / 1.5 - np.array([0, 0, 9.81]), atol=1e-12)


def test_dropin_planner_api():
    """Reads like the reference's own tests (tests/test_planner_controller_contract.py:52-87,
    tests/test_se3_mpc_with_mapper.py:9-42, tests/test_sitl_unit_tests.py:43-48)."""
    import dart_planner_b200 as dp
    planner = dp.SE3MPCPlanner()
    assert planner.config.prediction_horizon == 6 and planner.config["dt"] == 1 / 400
    assert float(planner.mass) == 1.5 and planner.hover_thrust == pytest.approx(14.715)
    state = dp.DroneState(timestamp=0.0, position=np.array([0.0, 0.0, 2.0]))
    traj = planner.plan_trajectory(state, np.array([10.0, 0.0, 5.0]))
    N = 6
    assert traj.positions.shape == (N, 3) and traj.thrusts.shape == (N,) and traj.attitudes.shape == (N, 3)
    assert (np.abs(traj.attitudes[:, :2]) < np.pi / 2).all() and (traj.thrusts > 0).all()
    g1 = load_golden("G1")
    np.testing.assert_allclose(planner.last_result["x"], g1["x"][0], atol=1e-9)
    assert planner.last_result["nit"] == 3 and planner.last_result["nfev"] == 5
    assert planner.last_result["fun"] == pytest.approx(5233.477019206591, rel=1e-12)
    assert planner.convergence_history == [True] and planner.is_plan_valid(traj)
    # goal hysteresis (:199): a goal moved by < 0.5 m is ignored
    planner.plan_trajectory(state, np.array([10.2, 0.0, 5.0]))
    np.testing.assert_array_equal(planner.goal_position, [10.0, 0.0, 5.0])
    planner.plan_trajectory(state, np.array([11.0, 0.0, 5.0]))
    np.testing.assert_array_equal(planner.goal_position, [11.0, 0.0, 5.0])
    # YAML-configured planner reproduces G2 (N=8, dt=0.1)
    p2 = dp.SE3MPCPlanner.from_yaml()
    sol = p2._solve_se3_mpc(state) if p2.set_goal(np.array([10.0, 0, 5.0])) is None else None
    np.testing.assert_allclose(p2.last_result["x"], load_golden("G2")["x"][0], atol=1e-9)
    assert set(sol) == {"positions", "velocities", "thrust_vectors", "accelerations", "attitudes",
                        "body_rates", "thrusts"}
    # update_plan with no goal -> emergency hover trajectory (:739-753)
    p3 = dp.SE3MPCPlanner()
    em = p3.update_plan(state, [{"position": [1, 1, 1], "radius": 0.5}])
    assert em.positions.shape == (6, 3) and len(p3.obstacles) == 1
    assert dp.planner.PlannerFactory.create("se3_mpc").__class__ is dp.SE3MPCPlanner


def test_host_buffer_entry_matches_device_entry():
    """dart_se3mpc_solve_batch_host (host pointers in/out; staged single-copy path for B <= 512,
    row copies above) returns exactly what the device-pointer entry computes."""
    import ctypes as C
    import dart_planner_b200 as dp
    from dart_planner_b200 import _cabi
    from dart_planner_b200.config import make_params
    L = _cabi.lib()
    N = 8
    cfg = dp.SE3MPCConfig(prediction_horizon=N, dt=0.1)
    params = make_params(cfg)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
    for B in (1, 5, 512, 513, 2000):
        p0, v0, goal = bench_inputs(40 + B, B, 1.0)
        first = dp.plan_batch(p0, v0, goal, cfg, to_host=True)
        hg = (np.arange(B) % 4 != 0).astype(np.uint8)
        for warm, use_hg in ((None, None), (first.x, hg)):
            ref = dp.plan_batch(p0, v0, goal, cfg, x_warm=warm, has_goal=use_hg, to_host=True)
            soa = lambda a: np.ascontiguousarray(np.asarray(a, np.float64).reshape(B, -1).T)  # noqa: E731
            i_p0, i_v0, i_goal = soa(p0), soa(v0), soa(goal)
            i_xw = None if warm is None else soa(warm)
            x = np.zeros((9 * N, B)); cost = np.zeros(B)
            nit = np.zeros(B, np.int32); nfev = np.zeros(B, np.int32); status = np.zeros(B, np.int32)
            acc = np.zeros((3 * N, B)); att = np.zeros((3 * N, B)); rates = np.zeros((3 * N, B))
            thr = np.zeros((N, B))
            rc = L.dart_se3mpc_solve_batch_host(C.byref(params), B, vp(i_p0), vp(i_v0), vp(i_goal),
                                                vp(use_hg), vp(i_xw), vp(x), vp(cost), vp(nit), vp(nfev),
                                                vp(status), vp(acc), vp(att), vp(rates), vp(thr))
            assert rc == 0
            np.testing.assert_array_equal(x.T, ref.x)
            np.testing.assert_array_equal(cost, ref.cost)
            np.testing.assert_array_equal(nit, ref.nit)
            np.testing.assert_array_equal(nfev, ref.nfev)
            np.testing.assert_array_equal(status, ref.status)
            np.testing.assert_array_equal(acc.T.reshape(B, N, 3), ref.accelerations)
            np.testing.assert_array_equal(att.T.reshape(B, N, 3), ref.attitudes)
            np.testing.assert_array_equal(rates.T.reshape(B, N, 3), ref.body_rates)
            np.testing.assert_array_equal(thr.T, ref.thrusts)
        # outputs are optional
        rc = L.dart_se3mpc_solve_batch_host(C.byref(params), B, vp(i_p0), vp(i_v0), vp(i_goal), None, None,
                                            vp(x), None, None, None, None, None, None, None, None)
        assert rc == 0


def test_inline_division_is_ieee_exact():
    """The solver's division (hardware reciprocal seed + Newton + residual correction, no range
    checks) returns the correctly rounded IEEE quotient over the operand ranges the solver sees
    (and far beyond): 2^[-200, 200], both signs, 2e8 pairs; zero dividends give zero."""
    import ctypes as C
    import torch
    from dart_planner_b200 import _cabi
    L = _cabi.lib()
    for emax, n in ((40, 100_000_000), (200, 100_000_000)):
        bad = torch.zeros(1, dtype=torch.int64, device="cuda")
        rc = L.dart_ddiv_selftest(n, 12345 + emax, emax, bad.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        assert int(bad.item()) == 0, f"{int(bad.item())} of {n} quotients differ from IEEE division (emax {emax})"


@pytest.mark.parametrize("N,m,gm,maxiter", [(1, 10, 0, 15), (2, 10, 0, 15), (64, 10, 0, 15), (8, 3, 1, 40),
                                            (8, 1, 1, 30), (6, 5, 1, 25), (33, 4, 1, 20)])
def test_gpu_edge_configurations(oracle_mod, N, m, gm, maxiter):
    """Horizon 1 (alpha = i/max(N-1,1)), the largest supported horizon (64: 32 lanes x 2 timesteps),
    and small correction memories with the exact gradient, so the ring of pairs wraps many times."""
    import dart_planner_b200 as dp
    from dart_planner_b200.config import make_params
    from dart_planner_b200.planner import BatchWorkspace
    B = 512
    p0, v0, goal = bench_inputs(100 + N + m, B, 1.5)
    tol = 5e-2 if gm == 0 else 1e-7
    cfg = dp.SE3MPCConfig(prediction_horizon=N, dt=0.1, max_iterations=maxiter, convergence_tolerance=tol)
    op = oracle_mod.make_params(horizon=N, dt=0.1, max_iterations=maxiter, convergence_tolerance=tol,
                                consistent_gradient=gm)
    op.max_corrections = m
    ref = oracle_mod.solve_batch(op, p0, v0, goal, nthreads=16)
    ws = BatchWorkspace(make_params(cfg, gradient_mode=gm, max_corrections=m), B, pinned=False)
    ws.set_inputs_device(p0, v0, goal)
    sol = ws.solve_device().numpy()
    if gm == 1:
        assert ref.nit.max() > m            # the ring wrapped
    _compare(sol, ref, min_counter_agreement=0.98 if gm == 1 else 1.0)


def test_gpu_empty_and_no_goal_batches():
    import dart_planner_b200 as dp
    from dart_planner_b200.config import make_params
    from dart_planner_b200.planner import solve_batch_tensors
    import torch
    cfg = dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)
    inp = torch.zeros((9, 32), dtype=torch.float64, device="cuda")
    sol = solve_batch_tensors(make_params(cfg), inp, 0)                   # B = 0: no launch, no error
    assert sol.B == 0 and sol.x.shape[0] == 0
    p0, v0, goal = bench_inputs(3, 64, 1.0)
    none = dp.plan_batch(p0, v0, goal, cfg, has_goal=np.zeros(64, np.uint8), to_host=True)
    # without a goal the position terms vanish: positions stay at p0, the solve still runs
    np.testing.assert_allclose(none.positions, np.repeat(p0[:, None, :], 8, axis=1), atol=0)
    assert (none.nit >= 1).all()


def test_pipelined_host_path_matches_device_path():
    """>= 65536 problems from host buffers go through the chunked two-stream pipeline (pitched
    H2D / solve / pitched D2H per 32768-problem chunk): same bits as the resident solve."""
    import dart_planner_b200 as dp
    from dart_planner_b200.config import make_params
    from dart_planner_b200.planner import BatchWorkspace, HostSolution
    B, N = 65536 + 32768 + 4096, 8          # three chunks, the last one partial
    p0, v0, goal = bench_inputs(71, B, 1.0)
    ws = BatchWorkspace(make_params(dp.SE3MPCConfig(prediction_horizon=N, dt=0.1)), B, pinned=True)
    assert ws.B == ws.ld
    ws.set_inputs_device(p0, v0, goal)
    ref = ws.solve_device().numpy()
    ws.out.zero_()
    host = ws.solve_host(p0, v0, goal)
    np.testing.assert_array_equal(host.x, ref.x)
    np.testing.assert_array_equal(host.cost, ref.cost)
    np.testing.assert_array_equal(host.nfev, ref.nfev)
    np.testing.assert_array_equal(host.status, ref.status)
    np.testing.assert_array_equal(host.body_rates, ref.body_rates)
    np.testing.assert_array_equal(host.thrusts, ref.thrusts)


def test_gpu_build_variants_agree(oracle_mod, monkeypatch):
    """The latency build, the 168-register throughput build, the 64-thread-block build and the
    general 9-slot instantiation (DART_SE3MPC_NO_COLD) solve the same batch to the same answer
    (separate compilations may contract differently: agreement to 1e-11, identical counters)."""
    import dart_planner_b200 as dp
    p0, v0, goal = bench_inputs(55, 4096, 1.0)
    cfg = dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)
    monkeypatch.delenv("DART_SE3MPC_VARIANT", raising=False)
    base = dp.plan_batch(p0, v0, goal, cfg, to_host=True)
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=8, dt=0.1), p0, v0, goal, nthreads=16)
    _compare(base, ref)
    for env in ({"DART_SE3MPC_VARIANT": "5"}, {"DART_SE3MPC_VARIANT": "6"}, {"DART_SE3MPC_NO_COLD": "1"},
                {"DART_SE3MPC_VARIANT": "5", "DART_SE3MPC_NO_COLD": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        got = dp.plan_batch(p0, v0, goal, cfg, to_host=True)
        for k in env:
            monkeypatch.delenv(k)
        np.testing.assert_allclose(got.x, base.x, rtol=0, atol=1e-11, err_msg=str(env))
        np.testing.assert_array_equal(got.nfev, base.nfev)
        np.testing.assert_array_equal(got.status, base.status)
        np.testing.assert_allclose(got.body_rates, base.body_rates, rtol=0, atol=1e-9)


@pytest.mark.parametrize("variant", [None, "5"], ids=["default-build", "throughput-build"])
def test_gpu_on_demand_breakpoints_and_step_bound(oracle_mod, monkeypatch, variant):
    """The regime in which the on-demand Cauchy breakpoints and the on-demand step bound of the
    line search are all needed (tests/test_core_emulation.py has the counts): the CUDA path must
    agree with the oracle there as it does where they are skipped."""
    import dart_planner_b200 as dp
    rng = np.random.default_rng(11)
    B, N = 1500, 8
    kw = dict(position_weight=3.0, velocity_weight=0.1, max_iterations=50, convergence_tolerance=1e-5)
    p0 = rng.uniform(-10, 10, (B, 3))
    v0 = rng.uniform(-2, 2, (B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=N, dt=0.1, **kw), p0, v0, goal, nthreads=16)
    if variant:
        monkeypatch.setenv("DART_SE3MPC_VARIANT", variant)
    sol = dp.plan_batch(p0, v0, goal, dp.SE3MPCConfig(prediction_horizon=N, dt=0.1, **kw), to_host=True)
    same = (sol.nit == ref.nit) & (sol.nfev == ref.nfev) & (sol.status == ref.status)
    assert same.mean() > 0.995, f"counter mismatches: {np.where(~same)[0][:10]}"
    assert (np.abs(sol.x - ref.x).max(axis=1)[same] < 1e-8).all()


@pytest.mark.parametrize("N", [5, 6])
def test_gpu_six_lane_groups_agree_with_the_eight_lane_builds(oracle_mod, monkeypatch, N):
    """Horizons 5 and 6 on 6-lane groups (five problems per warp, se3mpc_core.cuh SubWarp6;
    DART_SE3MPC_VARIANT=9 latency build / 10 throughput build -- measured slower than the 8-lane
    builds and therefore not the default): their reductions add in the order of the 8-lane butterfly
    with two empty lanes, so the solves agree with the default builds to contraction noise with
    identical counters, cold and warm, and with the oracle."""
    import dart_planner_b200 as dp
    p0, v0, goal = bench_inputs(61 + N, 4099, 1.5)          # ragged: the last warp holds 4 of 5 groups
    cfg = dp.SE3MPCConfig(prediction_horizon=N, dt=0.1)
    monkeypatch.delenv("DART_SE3MPC_VARIANT", raising=False)
    base = dp.plan_batch(p0, v0, goal, cfg, to_host=True)
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=N, dt=0.1), p0, v0, goal, nthreads=16)
    _compare(base, ref)
    xw = base.x.copy()
    xw[:, 6 * N:] += np.random.default_rng(3).normal(0, 0.3, xw[:, 6 * N:].shape)
    base_w = dp.plan_batch(p0, v0, goal, cfg, x_warm=xw, to_host=True)
    for variant in ("9", "10"):
        monkeypatch.setenv("DART_SE3MPC_VARIANT", variant)
        got = dp.plan_batch(p0, v0, goal, cfg, to_host=True)
        got_w = dp.plan_batch(p0, v0, goal, cfg, x_warm=xw, to_host=True)
        monkeypatch.delenv("DART_SE3MPC_VARIANT")
        for g, b in ((got, base), (got_w, base_w)):
            np.testing.assert_allclose(g.x, b.x, rtol=0, atol=1e-10, err_msg=variant)
            np.testing.assert_array_equal(g.nit, b.nit)
            np.testing.assert_array_equal(g.nfev, b.nfev)
            np.testing.assert_array_equal(g.status, b.status)
            np.testing.assert_allclose(g.attitudes, b.attitudes, rtol=0, atol=1e-9)
            np.testing.assert_allclose(g.body_rates, b.body_rates, rtol=0, atol=1e-7)
        _compare(got, ref)


def test_steps_in_flight_on_several_streams(oracle_mod):
    """A stream of planning steps with four in flight (one CUDA stream each) under
    `steps_in_flight`: the library picks the throughput build for the total load; every step's
    result equals the step solved alone (counters identical, x to contraction rounding) and the
    oracle's; the hint is process-wide and is reset on exit."""
    import ctypes as C
    import torch
    import dart_planner_b200 as dp
    from dart_planner_b200 import _cabi
    from dart_planner_b200.config import make_params
    from dart_planner_b200.planner import BatchWorkspace, steps_in_flight
    B, D, K = 4096, 4, 12
    params = make_params(dp.SE3MPCConfig(prediction_horizon=8, dt=0.1))
    L = _cabi.lib()
    info = [C.c_int32() for _ in range(5)]
    L.dart_se3mpc_kernel_info(C.byref(params), B, *[C.byref(i) for i in info])
    regs_alone = info[4].value
    steps = [bench_inputs(300 + k, B, 1.0) for k in range(K)]
    alone = [dp.plan_batch(*st, dp.SE3MPCConfig(prediction_horizon=8, dt=0.1), to_host=True) for st in steps]
    ws = [BatchWorkspace(params, B, pinned=False, outputs="all") for _ in range(K)]
    for w, st in zip(ws, steps):
        w.set_inputs_device(*st)
    streams = [torch.cuda.Stream() for _ in range(D)]
    torch.cuda.synchronize()
    with steps_in_flight(D * B):
        L.dart_se3mpc_kernel_info(C.byref(params), B, *[C.byref(i) for i in info])
        assert info[4].value <= 168 < regs_alone          # the register-capped throughput build
        sols = [w.solve_device(streams[k % D]) for k, w in enumerate(ws)]
        torch.cuda.synchronize()
    L.dart_se3mpc_kernel_info(C.byref(params), B, *[C.byref(i) for i in info])
    assert info[4].value == regs_alone
    for k in (0, 5, 11):
        got = sols[k].numpy()
        np.testing.assert_array_equal(got.nfev, alone[k].nfev)
        np.testing.assert_array_equal(got.status, alone[k].status)
        np.testing.assert_allclose(got.x, alone[k].x, rtol=0, atol=1e-11)
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=8, dt=0.1), *steps[7], nthreads=16)
    _compare(sols[7].numpy(), ref)
    assert L.dart_se3mpc_set_inflight_hint(-1) != 0


@pytest.mark.parametrize("near_fraction", [0.0005, 0.3, 0.6])    # used / ignored (too many) / ignored (list overflows)
def test_long_solves_first_schedule_changes_no_result(oracle_mod, monkeypatch, near_fraction):
    """Large cold batches in the throughput build are scheduled "long solves first": a scan kernel
    lists the problems that start within ~1 m of their goal (they end in the degenerate
    line-search regime, ~4x the work) and the first tickets after every block's first round serve
    that list.  Scheduling only: with the list (a few members: used; 30 % of the batch: ignored),
    without it (DART_SE3MPC_NO_PRIO), ragged batch, has_goal mask -- identical results, and the
    oracle's."""
    import dart_planner_b200 as dp
    rng = np.random.default_rng(808)
    B = 40000 + 7
    p0, v0, goal = bench_inputs(809, B, 0.0)
    near = rng.random(B) < near_fraction
    near[-300:] |= rng.random(300) < 0.05                 # some in the last rounds of the batch
    near[:9000:1500] = True                               # some inside the first rounds (solved there, not twice)
    goal[near] = p0[near] + rng.normal(0, 0.25, (int(near.sum()), 3))
    hg = (np.arange(B) % 11 != 0).astype(np.uint8)
    cfg = dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)
    monkeypatch.delenv("DART_SE3MPC_NO_PRIO", raising=False)
    with_list = dp.plan_batch(p0, v0, goal, cfg, has_goal=hg, to_host=True)
    monkeypatch.setenv("DART_SE3MPC_NO_PRIO", "1")
    without = dp.plan_batch(p0, v0, goal, cfg, has_goal=hg, to_host=True)
    monkeypatch.delenv("DART_SE3MPC_NO_PRIO")
    for k in ("x", "cost", "nit", "nfev", "status", "attitudes", "body_rates", "thrusts"):
        np.testing.assert_array_equal(getattr(with_list, k), getattr(without, k), err_msg=k)
    assert (with_list.nfev[near & (hg != 0)] > 8).any()      # the planted problems are long solves
    sub = np.r_[np.where(near)[0][:400], rng.integers(0, B, 2000)]
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=8, dt=0.1), p0[sub], v0[sub], goal[sub],
                                 has_goal=hg[sub], nthreads=16)
    relf = np.abs(with_list.cost[sub] - ref.cost) / np.maximum(np.abs(ref.cost), 1.0)
    same = (with_list.nit[sub] == ref.nit) & (with_list.nfev[sub] == ref.nfev) & (with_list.status[sub] == ref.status)
    assert same.mean() >= 0.97 and (relf[same] <= COST_RTOL).all()
    assert np.abs(with_list.x[sub] - ref.x)[same].max() <= CTRL_ATOL


def _compare_solutions(sol, ref, what):
    """Solution-level parity for configurations away from the reference's defaults.  With other
    weights / tighter tolerances many solves end in the degenerate line-search regime (status 2,
    up to ~45 evaluations) where the COUNTERS depend on rounding noise -- the oracle and this
    kernel's own host emulation disagree on up to 12 % of them while x agrees to 1e-11 -- so the
    solution is what is held to the north-star tolerance; counters only have to agree mostly."""
    relf = np.abs(sol.cost - ref.cost) / np.maximum(np.abs(ref.cost), 1.0)
    dx = np.abs(sol.x - ref.x).max(axis=1)
    ok = (relf <= COST_RTOL) & (dx <= CTRL_ATOL)
    same = (sol.nit == ref.nit) & (sol.nfev == ref.nfev) & (sol.status == ref.status)
    assert ok.mean() >= 0.99, f"{what}: {(~ok).sum()} of {ok.size} solutions out of tolerance"
    assert ok[same].all(), f"{what}: out of tolerance with equal counters"
    assert same.mean() >= 0.85, f"{what}: counters agree on {same.mean():.3f}"
    np.testing.assert_allclose(sol.attitudes[ok], ref.attitudes[ok], atol=1e-5)
    np.testing.assert_allclose(sol.thrusts[ok], ref.thrusts[ok], atol=2e-4)


def test_gpu_random_configurations(oracle_mod):
    """Randomised SE3MPCConfig / airframe parameters (weights, bounds, tilt, mass, horizon, dt,
    tolerance, iteration cap), warm and cold, against the oracle: the kernel reads every parameter
    the reference solve reads."""
    import dart_planner_b200 as dp
    rng = np.random.default_rng(2024)
    for trial in range(16):
        N = int(rng.choice([3, 5, 6, 8, 10, 16, 24]))
        kw = dict(max_velocity=float(rng.uniform(3, 15)), max_thrust=float(rng.uniform(18, 40)),
                  min_thrust=float(rng.uniform(0.5, 4)), max_tilt_angle=float(rng.uniform(0.3, 1.2)),
                  position_weight=float(rng.uniform(10, 300)), velocity_weight=float(rng.uniform(1, 30)),
                  acceleration_weight=float(rng.uniform(0.2, 5)), thrust_weight=float(rng.uniform(0.02, 1)),
                  max_iterations=int(rng.integers(2, 20)), convergence_tolerance=float(rng.choice([0.1, 0.05, 0.01])))
        dt = float(rng.choice([0.0025, 0.05, 0.1, 0.2]))
        mass = float(rng.uniform(0.6, 3.0))
        B = 384
        p0 = rng.uniform(-10, 10, (B, 3))
        v0 = rng.uniform(-3, 3, (B, 3))
        goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(2, 9, (B, 1))], axis=1)
        cfg = dp.SE3MPCConfig(prediction_horizon=N, dt=dt, **kw)
        op = oracle_mod.make_params(horizon=N, dt=dt, mass=mass, **kw)
        ref = oracle_mod.solve_batch(op, p0, v0, goal, nthreads=16)
        sol = dp.plan_batch(p0, v0, goal, cfg, mass=mass, to_host=True)
        _compare_solutions(sol, ref, f"trial {trial} cold {kw}")
        # warm start from the cold solution with a tilted thrust history
        xw = ref.x.copy()
        xw[:, 6 * N:] += rng.normal(0, 0.3, xw[:, 6 * N:].shape)
        ref_w = oracle_mod.solve_batch(op, p0 + 0.1, v0, goal, x_warm=xw, nthreads=16)
        sol_w = dp.plan_batch(p0 + 0.1, v0, goal, cfg, mass=mass, x_warm=xw, to_host=True)
        _compare_solutions(sol_w, ref_w, f"trial {trial} warm {kw}")


@pytest.mark.parametrize("N,B", [(8, 4096), (8, 1), (6, 777), (3, 130), (13, 512), (20, 300)])
def test_row_output_zero_copy_matches_soa_output(N, B):
    """dart_se3mpc_solve_batch_rows: the kernel reads pinned host inputs and writes one packed row
    per problem straight into pinned host memory (no copies).  Same bits as the SoA device path,
    in every lane configuration, cold / warm / has_goal / map check."""
    import dart_planner_b200 as dp
    from dart_planner_b200.config import make_params
    from dart_planner_b200.planner import BatchWorkspace, HostSolution
    fields = ("x", "cost", "nit", "nfev", "status", "task", "accelerations", "attitudes", "body_rates", "thrusts")
    p0, v0, goal = bench_inputs(100 + N, B, 2.0)
    ws = BatchWorkspace(make_params(dp.SE3MPCConfig(prediction_horizon=N, dt=0.1)), B, pinned=True)
    assert ws.rows_supported and ws.row_stride % 16 == 0 and ws.row_stride >= 19 * N + 4
    ws.set_inputs_device(p0, v0, goal)
    ws.stage_host_inputs(p0, v0, goal)
    ref = ws.solve_device().numpy()
    got = HostSolution.from_packed_rows(N, ws.solve_rows().numpy())
    for f in fields:
        np.testing.assert_array_equal(getattr(got, f), getattr(ref, f), err_msg=f)
    assert got.first_hit is None
    assert (ws.h_rows.numpy()[:B, 19 * N + 4:] == 0).all()              # padding is zeroed
    # solve_host takes the row path below the pipelining threshold
    host = ws.solve_host(p0, v0, goal)
    np.testing.assert_array_equal(host.x, ref.x)
    # warm start (tilted thrusts: 9-slot kernel) with a has_goal mask, plus the fused map check
    rng = np.random.default_rng(N)
    xprev = ref.x.copy()
    xprev[:, 6 * N:] += rng.normal(0, 0.5, xprev[:, 6 * N:].shape)
    hg = (np.arange(B) % 3 != 0).astype(np.uint8)
    grid = dp.DenseOccupancyGrid((128, 128, 128), (-64, -64, -64), 0.4)
    grid.add_obstacles(rng.uniform(-15, 15, (32, 3)), rng.uniform(0.8, 2.5, 32))
    ws.set_warm(xprev)
    ws.set_has_goal(hg)
    ws.set_map(grid, 1.0, 0.6)
    ref2 = ws.solve_device().numpy()
    ws.h_rows.zero_()
    got2 = HostSolution.from_packed_rows(N, ws.solve_rows().numpy())
    for f in fields + ("first_hit",):
        np.testing.assert_array_equal(getattr(got2, f), getattr(ref2, f), err_msg=f)
    if B > 100:
        assert (got2.first_hit >= 0).any() and (got2.first_hit == -1).any()


def test_row_output_limits():
    import ctypes as C
    import torch
    import dart_planner_b200 as dp
    from dart_planner_b200 import _cabi
    from dart_planner_b200.config import make_params
    from dart_planner_b200.planner import BatchWorkspace
    L = _cabi.lib()
    assert L.dart_se3mpc_row_stride(C.byref(make_params(dp.SE3MPCConfig(prediction_horizon=8))), 0) == 160
    assert L.dart_se3mpc_row_stride(C.byref(make_params(dp.SE3MPCConfig(prediction_horizon=22))), 0) == 432
    assert L.dart_se3mpc_row_stride(C.byref(make_params(dp.SE3MPCConfig(prediction_horizon=8))), 1) == 32
    assert L.dart_se3mpc_row_stride(C.byref(make_params(dp.SE3MPCConfig(prediction_horizon=8))), 2) == 80
    assert L.dart_se3mpc_row_stride(C.byref(make_params(dp.SE3MPCConfig(prediction_horizon=8))), 3) == 0
    # the row of a 40-step horizon does not fit the staging block: refused, SoA entries serve it
    pr = make_params(dp.SE3MPCConfig(prediction_horizon=40, dt=0.1))
    assert L.dart_se3mpc_row_stride(C.byref(pr), 0) == 0
    assert L.dart_se3mpc_row_stride(C.byref(pr), 1) == 128            # controls rows fit up to N = 64
    ws = BatchWorkspace(pr, 8, pinned=True)
    assert not ws.rows_supported
    with pytest.raises(RuntimeError):
        ws.solve_rows()
    p0, v0, goal = bench_inputs(3, 8, 1.0)
    assert ws.solve_host(p0, v0, goal).x.shape == (8, 360)               # falls back to the staged copies
    # stride not a multiple of 16 / too small / misaligned rows: bad argument
    pr = make_params(dp.SE3MPCConfig(prediction_horizon=8, dt=0.1))
    inp = torch.zeros((9, 32), dtype=torch.float64, device="cuda")
    rows = torch.zeros((32, 176), dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream

    def call(ptr, stride, kind=0, check_map=0):
        i = inp.data_ptr()
        return L.dart_se3mpc_solve_batch_rows(C.byref(pr), 8, 32, i, i + 768, i + 1536, None, None, None,
                                              ptr, stride, kind, None, 0.0, 0.0, check_map, s)
    assert call(rows.data_ptr(), 176) == 0                                # device rows, wider stride: fine
    assert call(rows.data_ptr(), 156) == -1
    assert call(rows.data_ptr(), 144) == -1
    assert call(rows.data_ptr() + 8, 160) == -1
    assert call(None, 160) == -1
    assert call(rows.data_ptr(), 32, kind=1) == 0                         # controls rows
    assert call(rows.data_ptr(), 16, kind=1) == -1
    assert call(rows.data_ptr(), 80, kind=2) == 0                         # solution rows
    assert call(rows.data_ptr(), 64, kind=2) == -1
    assert call(rows.data_ptr(), 160, kind=3) == -1
    assert call(rows.data_ptr(), 32, kind=1, check_map=1) == -2           # no map check in controls rows
    torch.cuda.synchronize()


@pytest.mark.parametrize("N,B", [(8, 4096), (6, 333), (4, 50), (13, 200), (40, 64)])
def test_controls_rows_match_the_full_solution(N, B):
    """DART_ROWS_CONTROLS: thrust vectors, cost and counters only, straight into pinned host memory
    (a fifth of the bytes of a full row); same bits as the resident solve, cold and warm."""
    import dart_planner_b200 as dp
    from dart_planner_b200.config import make_params
    from dart_planner_b200.planner import BatchWorkspace, HostSolution
    p0, v0, goal = bench_inputs(200 + N, B, 2.0)
    pr = make_params(dp.SE3MPCConfig(prediction_horizon=N, dt=0.1))
    full = BatchWorkspace(pr, B, pinned=False)
    full.set_inputs_device(p0, v0, goal)
    ref = full.solve_device().numpy()
    ws = BatchWorkspace(pr, B, pinned=True, outputs="controls")
    assert ws.rows_supported and ws.row_stride == (3 * N + 4 + 15) // 16 * 16
    ws.stage_host_inputs(p0, v0, goal)
    got = HostSolution.from_control_rows(N, ws.solve_rows().numpy())
    np.testing.assert_array_equal(got.thrust_vectors, ref.thrust_vectors)
    np.testing.assert_array_equal(got.cost, ref.cost)
    for f in ("nit", "nfev", "status", "task"):
        np.testing.assert_array_equal(getattr(got, f), getattr(ref, f), err_msg=f)
    assert (ws.h_rows.numpy()[:B, 3 * N + 1: 3 * N + 4].view(np.int32)[:, 4] == -2).all()
    xw = ref.x.copy()
    xw[:, 6 * N:] += np.random.default_rng(N).normal(0, 0.4, xw[:, 6 * N:].shape)
    full.set_warm(xw)
    ws.set_warm(xw)
    ref_w = full.solve_device().numpy()
    got_w = HostSolution.from_control_rows(N, ws.solve_rows().numpy())
    np.testing.assert_array_equal(got_w.thrust_vectors, ref_w.thrust_vectors)
    np.testing.assert_array_equal(got_w.nfev, ref_w.nfev)


@pytest.mark.parametrize("N,B", [(8, 4096), (6, 333), (4, 50), (13, 200), (40, 64)])
def test_solution_rows_match_the_full_solution(N, B):
    """DART_ROWS_SOLUTION: x, cost and counters (what scipy's minimize returns) straight into pinned
    host memory, half the bytes of a full row; the derived arrays are evaluated on the host from the
    thrust rows on first access (derive.py restates :582-654).  x / cost / counters: same bits as
    the resident solve; derived arrays: the tolerances of the fixture tests.  Cold, warm (tilted
    thrusts, zero-thrust steps) and with the fused map check."""
    import dart_planner_b200 as dp
    from dart_planner_b200.config import make_params
    from dart_planner_b200.planner import BatchWorkspace, HostSolution
    p0, v0, goal = bench_inputs(300 + N, B, 2.0)
    pr = make_params(dp.SE3MPCConfig(prediction_horizon=N, dt=0.1, min_thrust=0.0))
    full = BatchWorkspace(pr, B, pinned=False)
    full.set_inputs_device(p0, v0, goal)
    ws = BatchWorkspace(pr, B, pinned=True, outputs="solution")
    assert ws.rows_supported and ws.row_stride == (9 * N + 4 + 15) // 16 * 16
    assert ws.d2h_bytes_rows == ws.row_stride * 8 * B
    ws.stage_host_inputs(p0, v0, goal)

    def check(ref, got):
        np.testing.assert_array_equal(got.x, ref.x)
        np.testing.assert_array_equal(got.cost, ref.cost)
        for f in ("nit", "nfev", "status", "task"):
            np.testing.assert_array_equal(getattr(got, f), getattr(ref, f), err_msg=f)
        np.testing.assert_allclose(got.thrusts, ref.thrusts, rtol=0, atol=1e-12)
        np.testing.assert_allclose(got.accelerations, ref.accelerations, rtol=0, atol=1e-12)
        np.testing.assert_allclose(got.attitudes, ref.attitudes, rtol=0, atol=1e-12)
        np.testing.assert_allclose(got.body_rates, ref.body_rates, rtol=0, atol=1e-9)

    check(full.solve_device().numpy(), HostSolution.from_solution_rows(N, ws.solve_rows().numpy(), pr))
    assert (ws.h_rows.numpy()[:B, 9 * N + 1: 9 * N + 4].view(np.int32)[:, 4] == -2).all()
    assert (ws.h_rows.numpy()[:B, 9 * N + 4:] == 0).all()
    rng = np.random.default_rng(N)
    xw = full.solve_device().numpy().x.copy()
    xw[:, 6 * N:] += rng.normal(0, 0.8, xw[:, 6 * N:].shape)
    xw[::5, 6 * N + 3: 6 * N + 6] = 0.0           # a zero thrust that min_thrust = 0 keeps: invalid step
    grid = dp.DenseOccupancyGrid((128, 128, 128), (-64, -64, -64), 0.4)
    grid.add_obstacles(rng.uniform(-15, 15, (32, 3)), rng.uniform(0.8, 2.5, 32))
    for w in (full, ws):
        w.set_warm(xw)
        w.set_map(grid, 1.0, 0.6)
    ref_w = full.solve_device().numpy()
    got_w = ws.solve_host(p0, v0, goal)
    check(ref_w, got_w)
    np.testing.assert_array_equal(got_w.first_hit, ref_w.first_hit)


def test_gpu_chaotic_configuration_within_the_oracles_own_sensitivity(oracle_mod):
    """A configuration (found by an extended randomised sweep) where the reference algorithm is
    chaotic: the oracle disagrees with ITSELF on 60 % of the counters and 8 % of the solutions
    after moving every start position by one ulp.  The kernel must agree with the oracle as well
    as the oracle agrees with itself -- in the latency and in the throughput build."""
    import dart_planner_b200 as dp
    from conftest import CHAOTIC_CONFIG as cc, agreement, chaotic_inputs
    op = oracle_mod.make_params(horizon=cc["horizon"], dt=cc["dt"], mass=cc["mass"], **cc["kw"])
    cfg = dp.SE3MPCConfig(prediction_horizon=cc["horizon"], dt=cc["dt"], **cc["kw"])
    for B in (3000, 6000):          # one round of the latency build / the throughput build
        p0, v0, goal = chaotic_inputs(B)
        ref = oracle_mod.solve_batch(op, p0, v0, goal, nthreads=16)
        alt = oracle_mod.solve_batch(op, np.nextafter(p0, np.inf), v0, goal, nthreads=16)
        got = dp.plan_batch(p0, v0, goal, cfg, mass=cc["mass"], to_host=True)
        ok_self, same_self = agreement(alt, ref)
        ok_gpu, same_gpu = agreement(got, ref)
        assert ok_self < 0.97 and same_self < 0.7
        assert ok_gpu >= ok_self - 0.03 and same_gpu >= same_self - 0.06, (B, ok_gpu, ok_self, same_gpu, same_self)
