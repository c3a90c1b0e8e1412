"""the wire format and the goal producer either side of the batched solve, against
fixtures written by the UNMODIFIED reference classes (tools/gen_golden_host.py):
tests/golden/wire.npz (SecureSerializer messages) and tests/golden/mission.npz
(GlobalMissionPlanner.get_current_goal along 2 x 24 simulated flights)."""
import json

import numpy as np
import pytest

from conftest import load_golden


def _msgs(d):
    return [bytes(d[f"msg{i}"].tobytes()) for i in range(6)]


def test_reference_messages_verify_and_decode():
    from dart_planner_b200.wire import SignedEnvelope, trajectory_from_payload
    d = load_golden("wire")
    env = SignedEnvelope(secret_key=str(d["secret"]))
    for i, m in enumerate(_msgs(d)):
        got = env.deserialize(m, now=float(d["stamps"][i]) + 1.0)
        body = got["data"] if i == 4 else got
        np.testing.assert_array_equal(np.array([np.asarray(r) for r in body["positions"]]), d["P"][i])
        np.testing.assert_array_equal(np.asarray(body["timestamps"]), d["T"][i])
        if i == 3:
            assert body["velocities"] is None
        else:
            np.testing.assert_array_equal(np.array([np.asarray(r) for r in body["velocities"]]), d["V"][i])
        tr = trajectory_from_payload(body)
        assert tr.positions.shape == (8, 3) and tr.timestamps.shape == (8,)
        if i == 4:
            assert got["status"] == "success"
        if i == 5:
            assert got["n"] == 8


def test_serialize_reproduces_the_reference_bytes():
    """Same payload, timestamp and message id -> the same bytes as the reference serializer."""
    from dart_planner_b200.wire import SignedEnvelope
    d = load_golden("wire")
    env = SignedEnvelope(secret_key=str(d["secret"]))
    for i, m in enumerate(_msgs(d)):
        P, V, T = d["P"][i], d["V"][i], d["T"][i]
        payload = {"positions": P.tolist(), "velocities": V.tolist() if i != 3 else None, "timestamps": T.tolist()}
        if i == 4:
            payload = {"status": "success", "data": payload}
        if i == 5:
            payload = {"positions": P, "velocities": V, "timestamps": T, "n": np.int64(8)}
        mine = env.serialize(payload, timestamp=float(d["stamps"][i]), message_id=f"msg_{i + 1}_{int(d['pid'])}")
        assert mine == m, i


def test_tampered_expired_and_malformed_messages_are_refused():
    from dart_planner_b200.wire import SignedEnvelope, WireError
    d = load_golden("wire")
    env = SignedEnvelope(secret_key=str(d["secret"]))
    m, ts = _msgs(d)[0], float(d["stamps"][0])
    assert env.deserialize(m, now=ts + 299.0)
    with pytest.raises(WireError, match="too old"):
        env.deserialize(m, now=ts + 301.0)
    with pytest.raises(WireError, match="too old"):
        SignedEnvelope(secret_key=str(d["secret"]), message_ttl=5).deserialize(m, now=ts + 6.0)
    j = json.loads(m)
    j["data"]["positions"][0][0] += 1e-9
    with pytest.raises(WireError, match="signature"):
        env.deserialize(json.dumps(j).encode(), now=ts)
    with pytest.raises(WireError, match="signature"):
        SignedEnvelope(secret_key="another key").deserialize(m, now=ts)
    for bad in (b"not json", b"[1, 2]", json.dumps({"data": 1}).encode(), b"\xff\xfe"):
        with pytest.raises(WireError, match="Invalid message format"):
            env.deserialize(bad, now=ts)
    with pytest.raises(WireError):          # no key outside test mode (:54-55)
        import os
        old = {k: os.environ.pop(k, None) for k in ("DART_ZMQ_SECRET", "DART_ENVIRONMENT")}
        try:
            SignedEnvelope()
        finally:
            for k, v in old.items():
                if v is not None:
                    os.environ[k] = v
    assert SignedEnvelope(test_mode=True).secret_key          # test mode: a random key


def test_batched_trajectory_payloads_round_trip():
    """One message per drone of a batched solve: message ids count up, every message verifies,
    the decoded trajectory equals the solution rows, timestamps are t0 + k dt (:656-675)."""
    from dart_planner_b200.planner import HostSolution
    from dart_planner_b200.wire import SignedEnvelope, batch_trajectory_payloads, trajectory_from_payload
    rng = np.random.default_rng(0)
    B, N = 5, 8
    x = rng.normal(0, 3, (B, 9 * N))
    z = np.zeros(B, np.int32)
    sol = HostSolution(x=x, cost=np.zeros(B), nit=z, nfev=z, status=z, task=z, accelerations=None, attitudes=None,
                       body_rates=None, thrusts=None)
    env = SignedEnvelope(secret_key="k")
    t0 = np.array([10.0, 10.1, 10.2, 10.3, 10.4])
    payloads = batch_trajectory_payloads(sol, t0, 0.1, ids=[f"drone{b}" for b in range(B)])
    msgs = [env.serialize(p) for p in payloads]
    ids = [json.loads(m)["message_id"] for m in msgs]
    assert [i.split("_")[1] for i in ids] == ["1", "2", "3", "4", "5"]
    for b, m in enumerate(msgs):
        back = env.deserialize(m)
        tr = trajectory_from_payload(back)
        np.testing.assert_array_equal(tr.positions, sol.positions[b])
        np.testing.assert_array_equal(tr.velocities, sol.velocities[b])
        np.testing.assert_allclose(tr.timestamps, t0[b] + 0.1 * np.arange(N), rtol=0, atol=1e-12)
        assert back["drone_id"] == f"drone{b}"


@pytest.mark.parametrize("key", ["a", "b"])
def test_batched_goal_producer_matches_the_reference(key):
    """24 drones x 120 calls of get_current_goal: goals, phases and waypoint progress equal the
    reference planner's, drone by drone (takeoff, navigation with obstacle / doorway / landing-pad
    approaches, mapping, exploration spiral, landing, emergency on the ground)."""
    from dart_planner_b200.mission import BatchedMissionGoals, SemanticWaypoint
    d = load_golden("mission")
    wps = [SemanticWaypoint(p, str(lab)) for p, lab in zip(d[f"{key}_wp_pos"], d[f"{key}_wp_label"])]
    pos, goals = d[f"{key}_pos"], d[f"{key}_goals"]
    T, D, _ = pos.shape
    mg = BatchedMissionGoals(D, wps)
    mg.phase[:] = d[f"{key}_phase0"]
    seen = set()
    for t in range(T):
        g = mg.get_current_goals(pos[t], now=float(d["now0"]) + float(d["dnow"]) * t)
        np.testing.assert_allclose(g, goals[t], rtol=0, atol=1e-12, err_msg=f"step {t}")
        np.testing.assert_array_equal(mg.phase, d[f"{key}_phase_after"][t], err_msg=f"phase, step {t}")
        np.testing.assert_array_equal(mg.waypoint_index, d[f"{key}_wp_index_after"][t], err_msg=f"waypoint, step {t}")
        seen.update(mg.phase.tolist())
    assert {1, 3, 5} <= seen and (4 in seen or key == "a") and (2 in seen)


def test_goal_producer_without_waypoints_holds_position():
    from dart_planner_b200.mission import NAVIGATION, BatchedMissionGoals
    mg = BatchedMissionGoals(3)
    mg.phase[:] = NAVIGATION
    P = np.array([[1.0, 2.0, 3.0], [0.0, 0.0, 6.0], [4.0, 4.0, 4.0]])
    np.testing.assert_array_equal(mg.get_current_goals(P, now=0.5), P)       # :298-300
