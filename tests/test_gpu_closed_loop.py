"""BASELINE configs[4]: closed-loop receding-horizon simulation with resident state and real
warm starts, one launch per replanning step.  Oracle: the same loop on the
CPU (oracle solve with x_warm = previous solution, NumPy plant).  The reference never stores
`last_solution`, so this is an extension checked against our own restatement (self-oracle);
the warm-start construction (:294-327) and the plant model (:445-459) are the reference's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(seed, B):
    rng = np.random.default_rng(seed)
    p0 = rng.uniform(-10, 10, (B, 3))
    v0 = rng.uniform(-2, 2, (B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    return p0, v0, goal


def _plant(p, v, x, N, dt, mass=1.5, g=9.81):
    a = x[:, 6 * N: 6 * N + 3] / mass - np.array([0, 0, g])
    return p + v * dt + 0.5 * a * dt ** 2, v + a * dt


@pytest.mark.parametrize("N,dt,steps", [(8, 0.1, 40), (6, 0.05, 25)])
def test_closed_loop_matches_cpu_loop(oracle_mod, N, dt, steps):
    """Every replanning step against the CPU oracle started from the SAME state (the GPU's own
    p, v and previous solution), so a step's error is not carried into the next comparison --
    the free-running loops are compared separately below, over the first steps, because the
    reference controller is chaotic for part of the population (differences of 1e-10 grow
    ~1.5x per step)."""
    import dart_planner_b200 as dp
    from dart_planner_b200.closed_loop import ClosedLoopSim
    from dart_planner_b200.config import make_params
    B = 2048
    p0, v0, goal = _inputs(4, B)
    goal[:256] = p0[:256] + 0.02          # a near-goal population (degenerate line searches)
    sim = ClosedLoopSim(make_params(dp.SE3MPCConfig(prediction_horizon=N, dt=dt)), B, plant_dt=dt)
    sim.reset(p0, v0, goal)
    op = oracle_mod.make_params(horizon=N, dt=dt)
    pf, vf, xf = p0.copy(), v0.copy(), None          # free-running CPU loop
    for s in range(steps):
        pg, vg = sim.positions().cpu().numpy().copy(), sim.velocities().cpu().numpy().copy()
        xg = sim.solution().cpu().numpy().copy() if s > 0 else None
        sim.step()
        ref = oracle_mod.solve_batch(op, pg, vg, goal, x_warm=xg, nthreads=16)
        meta = sim.meta[:, :B].cpu().numpy()
        same = (meta[0] == ref.nit) & (meta[1] == ref.nfev) & (meta[2] == ref.status)
        assert same.mean() >= 0.99, f"step {s}: counters agree on {same.mean():.4f}"
        xs = sim.solution().cpu().numpy()
        ex = np.abs(xs - ref.x).max(axis=1)
        assert (ex[same] <= 1e-4).all(), f"step {s}: {(ex[same] > 1e-4).sum()} solutions beyond 1e-4"
        relf = np.abs(sim.cost[:B].cpu().numpy() - ref.cost) / np.maximum(np.abs(ref.cost), 1.0)
        assert (relf[same] <= 1e-5).all()
        # the plant step is exact arithmetic on the GPU's own first control
        p1, v1 = _plant(pg, vg, xs, N, dt)
        np.testing.assert_array_equal(sim.positions().cpu().numpy(), p1)
        np.testing.assert_array_equal(sim.velocities().cpu().numpy(), v1)
        if s < 10:
            rf = oracle_mod.solve_batch(op, pf, vf, goal, x_warm=xf, nthreads=16)
            pf, vf = _plant(pf, vf, rf.x, N, dt)
            xf = rf.x
            assert (np.abs(p1 - pf).max(axis=1) <= 1e-6).mean() >= 0.995, f"free-running step {s}"
    assert int(sim.nfev_total[:B].min()) >= steps and sim.steps_done == steps


@pytest.mark.parametrize("B", [777, 20000])
def test_closed_loop_first_step_is_the_plain_solve_plus_plant(B):
    """B = 20 000 is a large cold batch with drones that start within a metre of their goals: the
    "long solves first" schedule of plain solves must not be used by the launch that also moves
    the state in place (a drone solved in a priority round was stepped a second time)."""
    import dart_planner_b200 as dp
    from dart_planner_b200.closed_loop import ClosedLoopSim
    from dart_planner_b200.config import make_params
    N, dt = 8, 0.1
    p0, v0, goal = _inputs(9, B)
    goal[::200] = p0[::200] + np.random.default_rng(10).uniform(-0.5, 0.5, (len(p0[::200]), 3))   # a short list: it is used
    cfg = dp.SE3MPCConfig(prediction_horizon=N, dt=dt)
    sol = dp.plan_batch(p0, v0, goal, cfg, to_host=True)
    sim = ClosedLoopSim(make_params(cfg), B)
    sim.reset(p0, v0, goal)
    sim.step()
    np.testing.assert_array_equal(sim.solution().cpu().numpy(), sol.x)
    p1, v1 = _plant(p0, v0, sol.x, N, dt)
    np.testing.assert_array_equal(sim.positions().cpu().numpy(), p1)      # individually rounded ops
    np.testing.assert_array_equal(sim.velocities().cpu().numpy(), v1)
    # second step = warm-started solve from the advanced state
    warm = dp.plan_batch(p1, v1, goal, cfg, x_warm=sol.x, to_host=True)
    sim.step()
    np.testing.assert_array_equal(sim.solution().cpu().numpy(), warm.x)


def test_closed_loop_seven_slot_warm_path_is_identical_and_checked():
    """Warm starts whose lateral thrust is exactly zero run on the 7-slot kernel (warm=2): same
    bits as the general warm kernel; a problem that breaks the promise is refused (status 3)."""
    import dart_planner_b200 as dp
    from dart_planner_b200.closed_loop import ClosedLoopSim
    from dart_planner_b200.config import make_params
    B, N, dt = 3000, 8, 0.1
    p0, v0, goal = _inputs(17, B)
    params = make_params(dp.SE3MPCConfig(prediction_horizon=N, dt=dt))
    fast, gen = ClosedLoopSim(params, B, plant_dt=dt), ClosedLoopSim(params, B, plant_dt=dt)
    fast.reset(p0, v0, goal)
    gen.reset(p0, v0, goal)
    gen.lateral_thrust_is_zero = False           # general 9-slot warm kernel
    for s in range(12):
        fast.step()
        gen.step()
        gen.lateral_thrust_is_zero = False
    np.testing.assert_array_equal(fast.solution().cpu().numpy(), gen.solution().cpu().numpy())
    np.testing.assert_array_equal(fast.positions().cpu().numpy(), gen.positions().cpu().numpy())
    np.testing.assert_array_equal(fast.nfev_total.cpu().numpy(), gen.nfev_total.cpu().numpy())
    # break the promise for a few problems: tilt their stored thrust
    bad = [5, 77, 2999]
    before = fast.positions().cpu().numpy().copy()
    fast.x[6 * N + 3, bad] = 0.25                 # T_x of timestep 1
    fast.step()
    status = fast.meta[2, :B].cpu().numpy()
    assert (status[bad] == 3).all() and (np.delete(status, bad) != 3).all()
    assert np.isnan(fast.cost[:B].cpu().numpy()[bad]).all()
    np.testing.assert_array_equal(fast.positions().cpu().numpy()[bad], before[bad])   # state untouched


@pytest.mark.parametrize("B", [20000, 65536 + 77])
def test_run_in_sub_populations_changes_no_result(B):
    """`ClosedLoopSim.run` replans the population as independent sub-populations kept in flight on
    several streams (a drone's step k+1 depends on its own step k only): bit-identical to one
    launch per step for everybody, for every sub-population count."""
    import torch
    import dart_planner_b200 as dp
    from dart_planner_b200.config import make_params
    rng = np.random.default_rng(12)
    p0 = rng.uniform(-10, 10, (B, 3))
    v0 = rng.uniform(-2, 2, (B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    params = make_params(dp.SE3MPCConfig(prediction_horizon=8, dt=0.1))
    ref = dp.ClosedLoopSim(params, B, plant_dt=0.1)
    ref.reset(p0, v0, goal)
    ref.run(12, parts=1)
    torch.cuda.synchronize()
    for parts in (None, 2, 3, 4):
        sim = dp.ClosedLoopSim(params, B, plant_dt=0.1)
        sim.reset(p0, v0, goal)
        sim.run(5, parts=parts)
        sim.run(7, parts=parts)               # a second call continues from warm starts
        torch.cuda.synchronize()
        assert sim.steps_done == 12
        assert torch.equal(sim.positions(), ref.positions()) and torch.equal(sim.velocities(), ref.velocities())
        assert torch.equal(sim.solution(), ref.solution()) and torch.equal(sim.nfev_total, ref.nfev_total)
        assert torch.equal(sim.meta, ref.meta)
