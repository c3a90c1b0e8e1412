"""BASELINE configs[2]: 65 536 batched solves next to a 256^3 occupancy grid:
solve + the reference's post-hoc `is_trajectory_safe` on the solved positions, fused into the
solve kernel.  Integer outputs: exact against the stand-alone mapper kernel on the same
positions, and against the oracle (solve + stencil on the CPU) wherever the solves agree."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def config3_world(seed=2, n_spheres=64):
    rng = np.random.default_rng(seed)
    centers = rng.uniform(-20, 20, (n_spheres, 3))
    radii = rng.uniform(0.5, 2.0, n_spheres)
    return centers, radii


def config3_inputs(seed, B):
    rng = np.random.default_rng(seed)
    p0 = rng.uniform(-10, 10, (B, 3))
    v0 = rng.uniform(-2, 2, (B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    return p0, v0, goal


def test_fused_safety_check_matches_mapper_kernel_and_oracle(oracle_mod):
    import dart_planner_b200 as dp
    centers, radii = config3_world()
    grid = dp.DenseOccupancyGrid((256, 256, 256), (-128, -128, -128), 0.2)
    grid.add_obstacles(centers, radii)
    B = 65536
    p0, v0, goal = config3_inputs(2, B)
    cfg = dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)
    sol = dp.plan_batch(p0, v0, goal, cfg, grid=grid, safety_margin=1.5, collision_threshold=0.6)
    fused = sol.first_hit.cpu().numpy()
    # (1) the stand-alone batched is_trajectory_safe on the same device positions: identical
    alone = grid.trajectories_safe_soa(sol.out, B, 8, 1.5, 0.6)[:B].cpu().numpy()
    np.testing.assert_array_equal(fused, alone)
    assert (fused >= -1).all() and (fused < 8).all()
    assert 0.01 < (fused >= 0).mean() < 0.99            # the world is neither empty nor full
    # (2) the solve itself is unchanged by the fused check
    plain = dp.plan_batch(p0, v0, goal, cfg, to_host=True)
    host = sol.numpy()
    np.testing.assert_array_equal(host.x, plain.x)
    np.testing.assert_array_equal(host.nfev, plain.nfev)
    # (3) oracle: CPU solve + CPU stencil on a sample (the CPU stencil is a per-trajectory call)
    og = oracle_mod.DenseGrid((256, 256, 256), (-128, -128, -128), 0.2)
    for c, r in zip(centers, radii):
        og.add_sphere(c, r)
    np.testing.assert_array_equal(grid.occ.cpu().numpy(), og.occ)      # rasterisation is exact
    ns = 8192
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=8, dt=0.1), p0[:ns], v0[:ns], goal[:ns],
                                 nthreads=16)
    ref_hit = np.array([og.traj_safe(ref.positions[i], 1.5, 0.6) for i in range(ns)])
    same = (host.nit[:ns] == ref.nit) & (host.nfev[:ns] == ref.nfev)
    assert same.mean() == 1.0
    # positions agree to ~1e-10, so a voxel index can differ only for a point that close to a
    # voxel face: allow a handful, require the rest exact
    assert (fused[:ns] == ref_hit).mean() >= 0.9995
    # and the stencil evaluated by the CPU oracle on the GPU's own positions is exact
    own = np.array([og.traj_safe(host.positions[i], 1.5, 0.6) for i in range(2048)])
    np.testing.assert_array_equal(fused[:2048], own)


def test_fused_safety_other_horizons(oracle_mod):
    """Lane configurations with several timesteps per lane (N=13: 16 lanes, N=40: 32 lanes x 2)."""
    import dart_planner_b200 as dp
    centers, radii = config3_world(5, 24)
    grid = dp.DenseOccupancyGrid((128, 128, 128), (-64, -64, -64), 0.4)
    grid.add_obstacles(centers, radii)
    og = oracle_mod.DenseGrid((128, 128, 128), (-64, -64, -64), 0.4)
    for c, r in zip(centers, radii):
        og.add_sphere(c, r)
    for N in (4, 6, 13, 40):
        p0, v0, goal = config3_inputs(10 + N, 512)
        sol = dp.plan_batch(p0, v0, goal, dp.SE3MPCConfig(prediction_horizon=N, dt=0.1), grid=grid,
                            safety_margin=1.0, collision_threshold=0.6)
        host = sol.numpy()
        own = np.array([og.traj_safe(host.positions[i], 1.0, 0.6) for i in range(512)])
        np.testing.assert_array_equal(host.first_hit, own)


def test_grid_penalty_mode_matches_self_oracle(oracle_mod):
    """BASELINE configs[2] 'with mapper obstacle cost': gradient_mode 2 adds the occupancy-grid
    penalty to f and g inside the solve.  The reference computes no obstacle term,
    so this is an extension checked against our own CPU restatement (self-oracle)."""
    import dart_planner_b200 as dp
    centers, radii = config3_world()
    grid = dp.DenseOccupancyGrid((256, 256, 256), (-128, -128, -128), 0.2)
    grid.add_obstacles(centers, radii)
    og = oracle_mod.DenseGrid((256, 256, 256), (-128, -128, -128), 0.2)
    for c, r in zip(centers, radii):
        og.add_sphere(c, r)
    B = 16384
    p0, v0, goal = config3_inputs(2, B)
    cfg = dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)
    sol = dp.plan_batch(p0, v0, goal, cfg, grid=grid, safety_margin=1.5, collision_threshold=0.6,
                        obstacle_penalty=True)
    host = sol.numpy()
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=8, dt=0.1), p0, v0, goal, nthreads=16,
                                 grid=og, obstacle_weight=cfg.obstacle_weight, free_level=0.5)
    plain = dp.plan_batch(p0, v0, goal, cfg, to_host=True)
    changed = np.abs(host.x - plain.x).max(axis=1) > 1e-6
    assert changed.mean() > 0.05                      # the penalty is active for part of the batch
    same = (host.nit == ref.nit) & (host.nfev == ref.nfev) & (host.status == ref.status)
    assert same.mean() >= 0.995, f"counters agree on {same.mean():.4f}"
    relf = np.abs(host.cost - ref.cost) / np.maximum(np.abs(ref.cost), 1.0)
    dx = np.abs(host.x - ref.x).max(axis=1)
    assert (relf[same] <= 1e-5).all() and (dx[same] <= 1e-4).all()
    # the penalised solutions collide less often than the plain ones (sanity of the definition)
    hit_plain = grid.trajectories_safe_soa(dp.plan_batch(p0, v0, goal, cfg).out, B, 8, 1.5, 0.6)[:B].cpu().numpy()
    assert (host.first_hit >= 0).sum() <= (hit_plain >= 0).sum()
    # penalty mode without a grid is rejected
    with pytest.raises(ValueError):
        dp.plan_batch(p0[:4], v0[:4], goal[:4], cfg, obstacle_penalty=True)


@pytest.mark.parametrize("transport", ["gather", "host_block"])
@pytest.mark.parametrize("outputs", ["solution", "all"])
def test_sharded_solver_with_the_replicated_map(outputs, transport):
    """BASELINE configs[3] (Monte-Carlo initial states, one goal, sharded by problem index, map
    replicated per GPU): `ShardedSolver.set_map` runs the fused safety check against the rank's
    replica and the result rows carry it -- one process here (the multi-rank plumbing of the map
    is the gloo test of `replicate_map`); identical to the single-GPU fused check."""
    import dart_planner_b200 as dp
    from dart_planner_b200.config import make_params
    centers, radii = config3_world()
    grid = dp.replicate_map(dp.DenseOccupancyGrid((256, 256, 256), (-128, -128, -128), 0.2))
    grid.add_obstacles(centers, radii)
    B = 20011
    rng = np.random.default_rng(3)
    p0 = rng.normal((0, 0, 2), 1.0, (B, 3))
    v0 = rng.normal(0, 0.5, (B, 3))
    goal = np.tile([10.0, 0.0, 5.0], (B, 1))
    cfg = dp.SE3MPCConfig(prediction_horizon=8, dt=0.1)
    solver = dp.ShardedSolver(make_params(cfg), outputs=outputs, transport=transport)
    solver.set_map(grid, 1.5, 0.6)
    sol = solver.solve(p0, v0, goal, copy=True)
    one = dp.plan_batch(p0, v0, goal, cfg, grid=grid, safety_margin=1.5, collision_threshold=0.6)
    want = one.first_hit.cpu().numpy()
    assert sol.first_hit is not None
    np.testing.assert_array_equal(sol.first_hit, want)
    assert len(np.unique(want)) >= 4           # (this goal sits next to a sphere: the first colliding step varies)
    host = one.numpy()
    np.testing.assert_array_equal(sol.nfev, host.nfev)
    np.testing.assert_allclose(sol.x, host.x, rtol=0, atol=1e-11)     # rows: latency build; plan_batch: throughput build
    solver.set_map(None)
    again = solver.solve(p0, v0, goal, copy=True)
    assert again.first_hit is None
    np.testing.assert_array_equal(again.x, sol.x)
    if transport == "host_block":
        solver._block.close()
