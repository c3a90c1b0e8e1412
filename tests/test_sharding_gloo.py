"""world_size-2 (and 3, ragged) checks of the shard partition + single-gather plumbing on the
gloo backend.  The local solve is a stand-in built on the oracle (tests may use it); on the box
the same code runs over NCCL with the CUDA solve."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT  # noqa: F401


def test_shard_ranges_cover_and_balance():
    from dart_planner_b200.sharding import shard_counts, shard_range
    for B in (0, 1, 7, 4096, 65536, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            c = shard_counts(B, world)
            assert sum(c) == B and max(c) - min(c) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _oracle_solve_fn(params, p0, v0, goal):
    """Stand-in local solve: oracle result packed in the product's SoA block layout
    (rows [x 9N | cost | acc 3N | att 3N | rates 3N | thrust N], meta rows nit/nfev/status/task)."""
    import torch
    import oracle
    N = int(params.horizon)
    b = len(p0)
    out = np.zeros((19 * N + 1, b))
    meta = np.zeros((4, b), np.int32)
    if b:
        op = oracle.make_params(horizon=N, dt=float(params.dt))
        r = oracle.solve_batch(op, p0, v0, goal, nthreads=2)
        out[: 9 * N] = r.x.T
        out[9 * N] = r.cost
        out[9 * N + 1: 12 * N + 1] = r.accelerations.reshape(b, 3 * N).T
        out[12 * N + 1: 15 * N + 1] = r.attitudes.reshape(b, 3 * N).T
        out[15 * N + 1: 18 * N + 1] = r.body_rates.reshape(b, 3 * N).T
        out[18 * N + 1:] = r.thrusts.T
        meta[0], meta[1], meta[2] = r.nit, r.nfev, r.status
    return torch.from_numpy(out), torch.from_numpy(meta)


def _oracle_rows_fn(params, p0, v0, goal, outputs):
    """Stand-in local solve in the product's ROW layouts (include/dart_se3mpc.h: full / solution /
    controls rows, padded to a multiple of 16 doubles)."""
    import torch
    import oracle
    N = int(params.horizon)
    b = len(p0)
    payload = {"all": 19 * N + 4, "solution": 9 * N + 4, "controls": 3 * N + 4}[outputs]
    stride = (payload + 15) // 16 * 16
    rows = np.zeros((b, stride))
    if b:
        op = oracle.make_params(horizon=N, dt=float(params.dt))
        r = oracle.solve_batch(op, p0, v0, goal, nthreads=2)
        if outputs == "controls":
            rows[:, : 3 * N] = r.x[:, 6 * N:]
            at = 3 * N
        else:
            rows[:, : 9 * N] = r.x
            at = 9 * N
        rows[:, at] = r.cost
        if outputs == "all":
            rows[:, at + 1: at + 1 + 3 * N] = r.accelerations.reshape(b, 3 * N)
            rows[:, at + 1 + 3 * N: at + 1 + 6 * N] = r.attitudes.reshape(b, 3 * N)
            rows[:, at + 1 + 6 * N: at + 1 + 9 * N] = r.body_rates.reshape(b, 3 * N)
            rows[:, at + 1 + 9 * N: at + 1 + 10 * N] = r.thrusts
            at += 10 * N
        meta = rows[:, at + 1: at + 4].view(np.int32)
        meta[:, 0], meta[:, 1], meta[:, 2], meta[:, 4] = r.nit, r.nfev, r.status, -2
    return torch.from_numpy(rows)


def _rows_worker(rank, world, port, B, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dart_planner_b200.config import SE3MPCConfig, make_params
        from dart_planner_b200.sharding import ShardedSolver
        params = make_params(SE3MPCConfig(prediction_horizon=8, dt=0.1))
        rng = np.random.default_rng(3)
        p0 = rng.normal((0, 0, 2), 1.0, (B, 3))
        v0 = rng.normal(0, 0.5, (B, 3))
        goal = np.tile([10.0, 0.0, 5.0], (B, 1))
        res = {}
        for outputs in ("all", "solution", "controls"):
            solver = ShardedSolver(params, rows_fn=_oracle_rows_fn, outputs=outputs)
            sol = solver.solve(p0, v0, goal)
            sol_again = solver.solve(p0, v0, goal, copy=True)      # reuses the receive block
            if rank == 0:
                assert np.array_equal(sol.cost, sol_again.cost) and solver.last_timing["outputs"] == outputs
                if outputs == "controls":
                    res[outputs] = (sol.thrust_vectors.copy(), sol.cost.copy(), sol.nfev.copy())
                else:
                    res[outputs] = (sol.x.copy(), sol.cost.copy(), sol.nit.copy(), sol.nfev.copy(), sol.status.copy(),
                                    np.asarray(sol.thrusts).copy(), np.asarray(sol.body_rates).copy())
            else:
                assert sol is None
            # the same through the shared host block (no gather: every rank writes its slice)
            hb = ShardedSolver(params, rows_fn=_oracle_rows_fn, outputs=outputs, transport="host_block")
            s2 = hb.solve(p0, v0, goal)
            s3 = hb.solve(p0, v0, goal, copy=True)                 # reuses the shared block
            if rank == 0:
                key = "thrust_vectors" if outputs == "controls" else "x"
                assert np.array_equal(getattr(s2, key), res[outputs][0]) and np.array_equal(s2.cost, res[outputs][1])
                assert np.array_equal(getattr(s3, key), res[outputs][0]) and np.array_equal(s3.nfev, sol.nfev)
            else:
                assert s2 is None and s3 is None
            dist.barrier()
            hb._block.close()
        if rank == 0:
            q.put(res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _worker(rank, world, port, B, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dart_planner_b200.config import SE3MPCConfig, make_params
        from dart_planner_b200.sharding import ShardedSolver, shard_range
        params = make_params(SE3MPCConfig(prediction_horizon=8, dt=0.1))
        rng = np.random.default_rng(3)
        p0 = rng.normal((0, 0, 2), 1.0, (B, 3))
        v0 = rng.normal(0, 0.5, (B, 3))
        goal = np.tile([10.0, 0.0, 5.0], (B, 1))
        solver = ShardedSolver(params, solve_fn=_oracle_solve_fn)
        sol = solver.solve(p0, v0, goal)
        lo, hi = shard_range(B, world, rank)
        sol2 = solver.solve(p0[lo:hi], v0[lo:hi], goal[lo:hi], presliced=True, global_B=B)
        if rank == 0:
            assert np.array_equal(sol.x, sol2.x)
            q.put((sol.x, sol.cost, sol.nit, sol.nfev, sol.status, sol.thrusts, sol.body_rates))
        else:
            assert sol is None and sol2 is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,B", [(2, 64), (3, 50), (2, 1)])
def test_sharded_solve_gathers_in_problem_order(oracle_mod, world, B):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    p0 = rng.normal((0, 0, 2), 1.0, (B, 3))
    v0 = rng.normal(0, 0.5, (B, 3))
    goal = np.tile([10.0, 0.0, 5.0], (B, 1))
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=8, dt=0.1), p0, v0, goal, nthreads=2)
    x, cost, nit, nfev, status, thrusts, rates = got
    assert np.array_equal(x, ref.x) and np.array_equal(cost, ref.cost)
    assert np.array_equal(nit, ref.nit) and np.array_equal(nfev, ref.nfev) and np.array_equal(status, ref.status)
    assert np.array_equal(thrusts, ref.thrusts) and np.array_equal(rates, ref.body_rates)


@pytest.mark.parametrize("world,B", [(2, 64), (3, 50), (2, 1)])
def test_sharded_row_gather_all_row_kinds(oracle_mod, world, B):
    """The product's data path (one packed row per problem, one gather into one receive block,
    ragged slices, empty slices) for the three row kinds; the solution rows' derived arrays come
    from the host-side derivation."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    p0 = rng.normal((0, 0, 2), 1.0, (B, 3))
    v0 = rng.normal(0, 0.5, (B, 3))
    goal = np.tile([10.0, 0.0, 5.0], (B, 1))
    ref = oracle_mod.solve_batch(oracle_mod.make_params(horizon=8, dt=0.1), p0, v0, goal, nthreads=2)
    for kind in ("all", "solution"):
        x, cost, nit, nfev, status, thrusts, rates = got[kind]
        assert np.array_equal(x, ref.x) and np.array_equal(cost, ref.cost)
        assert np.array_equal(nit, ref.nit) and np.array_equal(nfev, ref.nfev) and np.array_equal(status, ref.status)
        np.testing.assert_allclose(thrusts, ref.thrusts, rtol=0, atol=1e-12)
        np.testing.assert_allclose(rates, ref.body_rates, rtol=0, atol=1e-9)
    tv, cost, nfev = got["controls"]
    assert np.array_equal(tv.reshape(B, -1), ref.x[:, 48:]) and np.array_equal(cost, ref.cost) and np.array_equal(nfev, ref.nfev)


def _map_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace
        from dart_planner_b200.sharding import replicate_map

        def host_grid(shape):       # what replicate_map reads of a DenseOccupancyGrid (the real one lives on a GPU)
            nx, ny, nz = shape
            return SimpleNamespace(nx=nx, ny=ny, nz=nz, origin_voxel=(-8, -6, -4), resolution=0.5, prob_prior=0.5,
                                   dtype="float32", occ=torch.full((nz, ny, nx), 0.5, dtype=torch.float32))
        grid = host_grid((16, 12, 8))                   # every rank constructs its own
        if rank == 0:       # the map the mapper built on the source rank
            grid.occ.copy_(torch.rand((8, 12, 16), generator=torch.Generator().manual_seed(5)).to(grid.occ.dtype))
        assert replicate_map(grid, src=0) is grid
        want = torch.rand((8, 12, 16), generator=torch.Generator().manual_seed(5)).to(grid.occ.dtype)
        ok = bool(torch.equal(grid.occ, want))
        # another geometry on one rank: every rank raises (nobody is left inside the broadcast)
        other = host_grid((16, 12, 8 if rank == 0 else 9))
        try:
            replicate_map(other, src=0)
            raised = False
        except ValueError:
            raised = True
        dist.barrier()
        q.put((rank, ok, raised))
    finally:
        dist.destroy_process_group()


def test_replicate_map_broadcasts_the_source_ranks_cells():
    """BASELINE configs[3]'s replicated map: one broadcast of the source rank's cells into every
    rank's own grid; a geometry mismatch fails on every rank together."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_map_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == [(0, True, True), (1, True, True)]
