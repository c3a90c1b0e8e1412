#!/usr/bin/env python
"""Generate tests/golden/wire.npz and tests/golden/mission.npz from the UNMODIFIED reference
(`communication/secure_serializer.py`, `planning/global_mission_planner.py`), imported from
/root/reference under the `pint` stand-in of tools/refshim.  Build container only.

    python tools/gen_golden_host.py

Wall-clock time and the process id are pinned through `time.time` / `os.getpid` stand-ins bound for
the duration of a call, so that the recorded messages are reproducible; the reference's code is
not altered."""
import logging
import os
import sys
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, "/root/reference/src")
os.environ["DART_ZMQ_SECRET"] = "golden-fixture-secret"
logging.disable(logging.CRITICAL)

from dart_planner.common.types import DroneState  # noqa: E402
from dart_planner.communication import secure_serializer as ss  # noqa: E402
from dart_planner.planning import global_mission_planner as gm  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def gen_wire():
    rng = np.random.default_rng(5)
    ser = ss.SecureSerializer(secret_key="golden-fixture-secret")
    msgs, stamps = [], []
    payloads = []
    N = 8
    for i in range(6):
        P = rng.uniform(-10, 10, (N, 3))
        V = rng.uniform(-3, 3, (N, 3))
        t = 1000.0 + 0.1 * np.arange(N) + i
        payloads.append((P, V, t))
        # the cloud node's trajectory answer (cloud/main_improved_threelayer.py:116-124)
        d = {"positions": P.tolist(), "velocities": V.tolist() if i != 3 else None, "timestamps": t.tolist()}
        if i == 4:
            d = {"status": "success", "data": d}            # wrapped the way zmq_server answers (:127-131)
        if i == 5:
            d = {"positions": P, "velocities": V, "timestamps": t, "n": np.int64(N)}   # raw ndarrays / NumPy scalars
        ts = 1.7e9 + 12.25 * i
        with mock.patch("time.time", return_value=ts), mock.patch("os.getpid", return_value=4242):
            msgs.append(ser.serialize(d))
        stamps.append(ts)
        with mock.patch("time.time", return_value=ts + 1.0):
            back = ser.deserialize(msgs[-1])
        assert isinstance(back, dict)
    np.savez_compressed(
        os.path.join(OUT, "wire.npz"), secret=np.array("golden-fixture-secret"), pid=np.int64(4242),
        stamps=np.array(stamps), P=np.array([p[0] for p in payloads]), V=np.array([p[1] for p in payloads]),
        T=np.array([p[2] for p in payloads]),
        **{f"msg{i}": np.frombuffer(m, dtype=np.uint8) for i, m in enumerate(msgs)})
    print("wrote wire.npz", [len(m) for m in msgs])


def gen_mission():
    """Two missions: "a" ends on a landing pad (approached from 3 m above, so its 2 m rule never
    fires: the drones hover over it, a quirk of :381-384); "b" has none and runs into LANDING."""
    missions = {
        "a": [([10.0, 0.0, 5.0], "safe_zone"), ([15.0, 10.0, 8.0], "observation_point"), ([12.0, 14.0, 6.0], "obstacle"),
              ([5.0, 12.0, 4.0], "doorway"), ([0.0, 0.0, 0.5], "landing_pad")],
        "b": [([6.0, 0.0, 5.0], "safe_zone"), ([8.0, 6.0, 7.0], "obstacle"), ([2.0, 7.0, 4.0], "doorway")],
    }
    names = ["takeoff", "exploration", "mapping", "navigation", "landing", "emergency"]
    out = {}
    for key, wps in missions.items():
        rng = np.random.default_rng(9 if key == "a" else 10)
        D, T = 24, 120
        pos = np.zeros((T, D, 3))
        goals = np.zeros((T, D, 3))
        phases = np.zeros((T, D), np.int32)
        wpi = np.zeros((T, D), np.int64)
        start = rng.uniform(-3, 3, (D, 3))
        start[:, 2] = rng.uniform(0.6, 2.0, D)
        start[::6, 2] = 0.2                                   # on the ground: EMERGENCY at the first replan
        speed = rng.uniform(0.5, 1.6, D)
        for d in range(D):
            pl = gm.GlobalMissionPlanner(gm.GlobalMissionConfig(use_neural_scene=False))
            pl.set_mission_waypoints([gm.SemanticWaypoint(np.array(p), lab, 0.1, 1) for p, lab in wps])
            if d % 8 == 5:
                pl.current_phase = gm.MissionPhase.EXPLORATION
            if d % 8 == 7:
                pl.current_phase = gm.MissionPhase.MAPPING
            p = start[d].copy()
            for t in range(T):
                now = 100.0 + 0.4 * t                          # replans fall on some steps, not all
                st = DroneState(timestamp=now, position=p.copy())
                # the reference's "simulated" uncertain regions are drawn with np.random (:433-441):
                # pinned off (no new region), so the exploration goal is the deterministic spiral
                with mock.patch("time.time", return_value=now), mock.patch("numpy.random.random", return_value=1.0):
                    g = np.asarray(pl.get_current_goal(st), dtype=float)
                pos[t, d], goals[t, d] = p, g
                phases[t, d] = names.index(pl.current_phase.value)
                wpi[t, d] = pl.current_waypoint_index
                step = g - p                                    # fly towards the goal, a bit noisy
                n = np.linalg.norm(step)
                if n > 1e-9:
                    p = p + step / n * min(n, speed[d]) + rng.normal(0, 0.02, 3)
        out.update({f"{key}_wp_pos": np.array([w[0] for w in wps]), f"{key}_wp_label": np.array([w[1] for w in wps]),
                    f"{key}_pos": pos, f"{key}_goals": goals, f"{key}_phase_after": phases, f"{key}_wp_index_after": wpi,
                    f"{key}_phase0": np.array([1 if d % 8 == 5 else (2 if d % 8 == 7 else 0) for d in range(D)], np.int32)})
        print(f"mission {key}: final phases", np.bincount(phases[-1], minlength=6).tolist(),
              "phases seen", np.bincount(phases.ravel(), minlength=6).tolist())
    np.savez_compressed(os.path.join(OUT, "mission.npz"), now0=np.float64(100.0), dnow=np.float64(0.4), **out)
    print("wrote mission.npz")


def gen_local_grid():
    """get_local_occupancy_grid (:221-248) and the cloud node's occupied-point -> sphere selection
    (cloud/main_improved_threelayer.py:381-398) on a map built by add_obstacle + two update_map
    scans of the reference mapper."""
    from dart_planner.perception.explicit_geometric_mapper import ExplicitGeometricMapper, SensorObservation
    rng = np.random.default_rng(21)
    mp = ExplicitGeometricMapper(resolution=0.5, max_range=40.0)
    cs = np.array([[4.0, 1.0, 2.5], [-3.0, -2.0, 1.0], [1.0, 5.0, 4.0]])
    rs = np.array([1.5, 1.0, 2.0])
    for c, r in zip(cs, rs):
        mp.add_obstacle(c, float(r))
    R = 400
    pos = np.tile([0.2, 0.3, 2.0], (R, 1)) + rng.normal(0, 0.05, (R, 3))
    ang = rng.uniform(0, 2 * np.pi, R)
    dirs = np.stack([np.cos(ang), np.sin(ang), rng.normal(0, 0.2, R)], axis=1)
    hit = rng.uniform(2.0, 20.0, R)
    hit[rng.random(R) < 0.7] = np.nan
    obs = [SensorObservation(position=pos[i], direction=dirs[i], hit_distance=None if np.isnan(hit[i]) else float(hit[i]),
                             max_range=40.0, timestamp=0.0) for i in range(R)]
    mp.update_map(obs[:200])
    mp.update_map(obs[200:])
    center, size = np.array([1.3, -0.7, 2.1]), 15.0
    grid, occ = mp.get_local_occupancy_grid(center, size)
    pts = grid[occ > 0.6]
    step = max(1, pts.shape[0] // 20)
    np.savez_compressed(os.path.join(OUT, "local_grid.npz"), res=np.float64(0.5), max_range=np.float64(40.0),
                        sph_c=cs, sph_r=rs, pos=pos, dir=dirs, hit=hit, center=center, size=np.float64(size),
                        occ=occ, grid_corner=grid[0, 0, 0], grid_last=grid[-1, -1, -1], grid_sample=grid[3, 7, 11],
                        n_occupied=np.int64(pts.shape[0]), spheres=pts[::step])
    print("wrote local_grid.npz", occ.shape, "occupied", pts.shape[0], "spheres", pts[::step].shape[0],
          "values", np.unique(np.round(occ, 6))[:8])


if __name__ == "__main__":
    gen_wire()
    gen_mission()
    gen_local_grid()
