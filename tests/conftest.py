"""Shared fixtures.  `gpu` marks tests that need a real B200 (run by the driver with -m gpu)."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


SOLVER_FIXTURES = sorted(
    os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz"))
    if os.path.basename(f)[:-4] not in ("extract", "mapper", "update_map", "wire", "mission", "local_grid"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_cfg(d):
    """SE3MPCConfig overrides stored with a fixture (keys cfg_*)."""
    cfg = {k[4:]: float(d[k]) for k in d.files if k.startswith("cfg_")}
    if "max_iterations" in cfg:
        cfg["max_iterations"] = int(cfg["max_iterations"])
    return cfg


# Tolerances of BASELINE.json north_star (fp64 parity mode)
COST_RTOL = 1e-5
CTRL_ATOL = 1e-4


def assert_solution_parity(got, d, what=""):
    """got: object with x (B,9N), cost, nit, nfev, status, attitudes, body_rates, thrusts, accelerations."""
    fun = d["fun"]
    # non-finite fixtures (nonfinite_*.npz): NaNs must sit in exactly the same places
    x_got, cost_got = np.asarray(got.x), np.asarray(got.cost)
    assert np.array_equal(np.isnan(x_got), np.isnan(d["x"])), f"{what}: NaN pattern of x differs"
    assert np.array_equal(np.isnan(cost_got), np.isnan(fun)), f"{what}: NaN pattern of the cost differs"
    fun = np.nan_to_num(fun, nan=0.0)
    relf = np.abs(np.nan_to_num(cost_got, nan=0.0) - fun) / np.maximum(np.abs(fun), 1.0)
    dx = np.abs(np.nan_to_num(x_got, nan=0.0) - np.nan_to_num(d["x"], nan=0.0)).max(axis=1)
    bad = np.where((relf > COST_RTOL) | (dx > CTRL_ATOL))[0]
    assert bad.size == 0, f"{what}: {bad.size} problems out of tolerance, first {bad[:5]}, relf {relf[bad[:5]]}, dx {dx[bad[:5]]}"
    for k in ("nit", "nfev", "status"):
        mism = np.where(np.asarray(getattr(got, k)) != d[k])[0]
        assert mism.size == 0, f"{what}: {k} differs on {mism.size} problems, first {mism[:5]}: got {np.asarray(getattr(got, k))[mism[:5]]} want {d[k][mism[:5]]}"
    np.testing.assert_allclose(got.accelerations, d["accelerations"], atol=1e-4 / 1.0, rtol=0)   # equal_nan
    np.testing.assert_allclose(got.thrusts, d["thrusts"], atol=1e-4, rtol=0)
    np.testing.assert_allclose(got.attitudes, d["attitudes"], atol=1e-6, rtol=0)
    # body rates are finite differences of R over dt: scale the tolerance by 1/dt
    np.testing.assert_allclose(got.body_rates, d["body_rates"], atol=1e-4 / float(d["dt"]) * 1e-2 + 1e-6, rtol=1e-6)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


# A configuration from the extended randomised sweep in which the reference algorithm itself is
# chaotic (tight tolerance, dt = 2.5 ms, 29 iterations allowed: most solves end in the degenerate
# line-search regime its inconsistent gradient creates).  Used by the sensitivity-normalised
# parity tests: parity there can only be asked for up to the oracle's own sensitivity to a
# ONE-ulp change of the inputs.
CHAOTIC_CONFIG = dict(
    horizon=6, dt=0.0025, mass=0.8339711942780523,
    kw=dict(max_velocity=9.63584924528805, max_thrust=18.053020479408932, min_thrust=1.7315781576030407,
            max_tilt_angle=0.7194424921567795, position_weight=44.425990443591985,
            velocity_weight=1.3704288337127277, acceleration_weight=0.6909859549195232,
            thrust_weight=0.8006471784609364, max_iterations=29, convergence_tolerance=0.01))


def chaotic_inputs(B, seed=31):
    rng = np.random.default_rng(seed)
    p0 = rng.uniform(-10, 10, (B, 3))
    v0 = rng.uniform(-3, 3, (B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(2, 9, (B, 1))], axis=1)
    return p0, v0, goal


def agreement(got, ref):
    """(fraction within the north-star tolerance, fraction with equal counters)."""
    relf = np.abs(got.cost - ref.cost) / np.maximum(np.abs(ref.cost), 1.0)
    dx = np.abs(got.x - ref.x).max(axis=1)
    ok = (relf <= COST_RTOL) & (dx <= CTRL_ATOL)
    same = (got.nit == ref.nit) & (got.nfev == ref.nfev) & (got.status == ref.status)
    return float(ok.mean()), float(same.mean())
