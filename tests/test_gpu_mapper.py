"""Mapper kernels against fixtures produced by the reference's ExplicitGeometricMapper and
against the oracle at scale -- exact integer parity."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_trace_ray_fixtures():
    import dart_planner_b200 as dp
    d = load_golden("mapper")
    g = dp.DenseOccupancyGrid((8, 8, 8), (0, 0, 0), 0.5)
    v = g._trace_ray([0, 0, 0], np.array([1, 1, 0]) / np.sqrt(2), 5.0)
    assert v == [tuple(r) for r in d["kat1"].tolist()]
    assert all(sum(abs(a - b) for a, b in zip(p, q)) == 1 for p, q in zip(v, v[1:]))
    v = g._trace_ray([0.3, -0.2, 1.1], [-1.0, 2.0, 0.5], 3.0)
    assert v == [tuple(r) for r in d["kat2"].tolist()]
    g2 = dp.DenseOccupancyGrid((8, 8, 8), (0, 0, 0), float(d["ray_res"]))
    count, vox = g2.trace_rays(d["ray_start"], d["ray_dir"], d["ray_dist"], max_vox=int(d["ray_len"].max()))
    count = count.cpu().numpy(); vox = vox.cpu().numpy()
    np.testing.assert_array_equal(count, d["ray_len"])
    off = 0
    for i, L in enumerate(d["ray_len"]):
        np.testing.assert_array_equal(vox[:L, :, i], d["ray_vox"][off:off + L])
        off += L


def test_sphere_query_safety_fixtures():
    import dart_planner_b200 as dp
    d = load_golden("mapper")
    g = dp.DenseOccupancyGrid((256, 256, 256), (-128, -128, -128), 0.2)
    g.add_obstacle([15.0, 5.0, 5.0], 2.0)
    assert int((g.occ > 0.6).sum()) == int(d["kat3_nvox"]) == 4163
    np.testing.assert_allclose(g.query_occupancy_batch(d["kat3_q"]).cpu().numpy(), d["kat3_occ"], atol=1e-7)
    assert g.is_trajectory_safe(d["kat3_traj"], 1.5, 0.6) == (False, 4)
    g = dp.DenseOccupancyGrid((128, 128, 128), (-64, -64, -64), 0.2)
    g.add_obstacles(d["sph_c"], d["sph_r"])
    np.testing.assert_allclose(g.query_occupancy_batch(d["sph_q"]).cpu().numpy(), d["sph_occ"], atol=1e-7)
    idx = g.are_trajectories_safe(d["sph_traj"], 1.5, 0.6).cpu().numpy()
    np.testing.assert_array_equal(idx, d["sph_idx"])
    np.testing.assert_array_equal((idx < 0).astype(np.int32), d["sph_safe"])


def test_config3_map_against_oracle(oracle_mod):
    """BASELINE configs[2] map: 256^3 @ 0.2 m, 64 spheres (seed 2); 65 536 random queries, rays
    and 8-point trajectories against the oracle."""
    import dart_planner_b200 as dp
    rng = np.random.default_rng(2)
    cs, rs = rng.uniform(-20, 20, (64, 3)), rng.uniform(0.5, 2.0, 64)
    g = dp.DenseOccupancyGrid((256, 256, 256), (-128, -128, -128), 0.2)
    g.add_obstacles(cs, rs)
    og = oracle_mod.DenseGrid((256, 256, 256), (-128, -128, -128), 0.2)
    for c, r in zip(cs, rs):
        og.add_sphere(c, float(r))
    np.testing.assert_array_equal(g.occ.cpu().numpy(), og.occ)
    q = rng.uniform(-27, 27, (4096, 3))                      # includes out-of-grid points
    np.testing.assert_array_equal(g.query_occupancy_batch(q).cpu().numpy(), og.query(q))
    trajs = rng.uniform(-20, 20, (2048, 1, 3)) + np.cumsum(rng.normal(0, 0.6, (2048, 8, 3)), axis=1)
    idx = g.are_trajectories_safe(trajs, 1.5, 0.6).cpu().numpy()
    want = np.array([og.traj_safe(t, 1.5, 0.6) for t in trajs])
    np.testing.assert_array_equal(idx, want)
    assert (want >= 0).any() and (want < 0).any()
    starts = rng.uniform(-20, 20, (2048, 3)); dirs = rng.normal(0, 1, (2048, 3)); dist = rng.uniform(0.1, 20, 2048)
    count, vox = g.trace_rays(starts, dirs, dist, max_vox=256)
    count, vox = count.cpu().numpy(), vox.cpu().numpy()
    for i in range(0, 2048, 8):
        v, n = oracle_mod.trace_ray(0.2, starts[i], dirs[i], dist[i], max_vox=256)
        assert n == count[i]
        np.testing.assert_array_equal(vox[:n, :, i], v)


def test_update_map_matches_reference_fixture_and_oracle(oracle_mod):
    """Batched update_map (:100-152): two scans, against the reference mapper's resulting dict
    (float32 grid storage -> 1e-6) and against the sequential oracle at LiDAR scale."""
    import dart_planner_b200 as dp
    d = load_golden("update_map")
    keys, probs = d["keys"], d["probs"]
    lo = keys.min(0) - 1
    shape = tuple(int(v) for v in (keys.max(0) - lo + 2))
    g = dp.DenseOccupancyGrid(shape, tuple(int(v) for v in lo), float(d["res"]),
                              max_range=float(d["mapper_max_range"]))
    s = int(d["split"])
    r1 = g.update_map(d["pos"][:s], d["dir"][:s], d["hit"][:s], d["obs_max_range"][:s])
    r2 = g.update_map(d["pos"][s:], d["dir"][s:], d["hit"][s:], d["obs_max_range"][s:])
    assert [r1["updated_voxels"], r2["updated_voxels"]] == d["updated"].tolist()   # exact visit counts
    occ = g.occ.cpu().numpy()
    idx = keys - lo
    np.testing.assert_allclose(occ[idx[:, 2], idx[:, 1], idx[:, 0]], probs, rtol=0, atol=1e-6)
    touched = occ != np.float32(0.5)
    assert int(touched.sum()) == len(keys)                        # exactly the reference's voxel set
    assert int(g._counts.abs().sum().item()) == 0                  # scratch left zeroed
    # scale: 200k rays from 8 sensors on a 256^3 grid, against the sequential oracle
    rng = np.random.default_rng(3)
    R = 200_000
    sensors = rng.uniform(-10, 10, (8, 3))
    pos = sensors[rng.integers(0, 8, R)]
    dirs = rng.normal(0, 1, (R, 3))
    hit = rng.uniform(0.5, 30.0, R)
    hit[rng.random(R) < 0.25] = np.nan
    g2 = dp.DenseOccupancyGrid((256, 256, 256), (-128, -128, -128), 0.2, max_range=25.0)
    out = g2.update_map(pos, dirs, hit, 30.0)
    og = oracle_mod.DenseGrid((256, 256, 256), (-128, -128, -128), 0.2)
    occ64 = np.full(og.occ.shape, 0.5)
    upd = oracle_mod.update_map(og, occ64, pos, dirs, hit, np.full(R, 30.0), 25.0)
    assert out["updated_voxels"] == upd and out["observations_processed"] == R
    got = g2.occ.cpu().numpy()
    np.testing.assert_array_equal(got != np.float32(0.5), occ64 != 0.5)      # same voxels touched
    np.testing.assert_allclose(got, occ64, rtol=0, atol=1e-6)
