"""Mapper half of the oracle against fixtures produced by the reference's
ExplicitGeometricMapper (perception/explicit_geometric_mapper.py) -- exact integer parity."""
import numpy as np

from conftest import load_golden


def test_trace_ray_kats(oracle_mod):
    d = load_golden("mapper")
    v, n = oracle_mod.trace_ray(0.5, [0, 0, 0], np.array([1, 1, 0]) / np.sqrt(2), 5.0)
    np.testing.assert_array_equal(v, d["kat1"])
    assert n == 15
    # reference tests/test_mapper_trace_ray.py:6-16 : 6-connected
    assert (np.abs(np.diff(v, axis=0)).sum(axis=1) == 1).all()
    v, n = oracle_mod.trace_ray(0.5, [0.3, -0.2, 1.1], [-1.0, 2.0, 0.5], 3.0)
    np.testing.assert_array_equal(v, d["kat2"])


def test_trace_ray_random(oracle_mod):
    d = load_golden("mapper")
    off = 0
    for i in range(len(d["ray_len"])):
        v, n = oracle_mod.trace_ray(float(d["ray_res"]), d["ray_start"][i], d["ray_dir"][i],
                                    float(d["ray_dist"][i]))
        L = int(d["ray_len"][i])
        assert n == L
        np.testing.assert_array_equal(v, d["ray_vox"][off:off + L])
        off += L


def test_sphere_query_and_safety(oracle_mod):
    d = load_golden("mapper")
    g = oracle_mod.DenseGrid((256, 256, 256), (-128, -128, -128), 0.2)
    assert g.add_sphere([15.0, 5.0, 5.0], 2.0) == int(d["kat3_nvox"]) == 4163
    np.testing.assert_allclose(g.query(d["kat3_q"]), d["kat3_occ"], atol=1e-7)
    idx = g.traj_safe(d["kat3_traj"], 1.5, 0.6)
    assert (int(idx < 0), idx) == tuple(d["kat3_safe"]) == (0, 4)
    g = oracle_mod.DenseGrid((128, 128, 128), (-64, -64, -64), 0.2)
    for c, r in zip(d["sph_c"], d["sph_r"]):
        g.add_sphere(c, float(r))
    np.testing.assert_allclose(g.query(d["sph_q"]), d["sph_occ"], atol=1e-7)
    for t, s, i in zip(d["sph_traj"], d["sph_safe"], d["sph_idx"]):
        idx = g.traj_safe(t, 1.5, 0.6)
        assert idx == i and int(idx < 0) == s


def test_bayes(oracle_mod):
    d = load_golden("mapper")
    p, seq = 0.5, []
    for h in [1, 1, 0, 1, 0, 0, 0]:
        p = oracle_mod.bayes(p, h)
        seq.append(p)
    np.testing.assert_allclose(seq, d["kat4"], atol=1e-15)
    np.testing.assert_allclose(seq[:3], [0.7, 0.844828, 0.890909], atol=1e-6)


def test_oracle_update_map_matches_reference_fixture(oracle_mod):
    """update_map (:100-152): two scans through the oracle vs the reference mapper's dict."""
    d = load_golden("update_map")
    keys, probs, counts = d["keys"], d["probs"], d["counts"]
    lo = keys.min(0) - 1
    shape = tuple(int(v) for v in (keys.max(0) - lo + 2))
    g = oracle_mod.DenseGrid(shape, tuple(int(v) for v in lo), float(d["res"]))
    occ = np.full(g.occ.shape, 0.5)
    cnt = np.zeros(g.occ.shape, np.int32)
    s = int(d["split"])
    u1 = oracle_mod.update_map(g, occ, d["pos"][:s], d["dir"][:s], d["hit"][:s], d["obs_max_range"][:s],
                               float(d["mapper_max_range"]), cnt)
    u2 = oracle_mod.update_map(g, occ, d["pos"][s:], d["dir"][s:], d["hit"][s:], d["obs_max_range"][s:],
                               float(d["mapper_max_range"]), cnt)
    assert [u1, u2] == d["updated"].tolist()
    idx = keys - lo
    np.testing.assert_allclose(occ[idx[:, 2], idx[:, 1], idx[:, 0]], probs, rtol=0, atol=1e-15)
    np.testing.assert_array_equal(cnt[idx[:, 2], idx[:, 1], idx[:, 0]], counts)
    assert int((cnt > 0).sum()) == len(keys)                       # no other voxel was touched
