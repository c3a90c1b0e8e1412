"""Wire format around the solve path: the reference's HMAC-signed JSON envelope
(src/dart_planner/communication/secure_serializer.py:92-169) and the trajectory message a cloud
node answers an edge drone with (src/dart_planner/cloud/main_improved_threelayer.py:116-124), for
ONE trajectory or for a whole batch of solves.

Envelope (a JSON object, UTF-8): ``{"data": <payload>, "signature": <hex>, "timestamp": <float>,
"message_id": "msg_<counter>_<pid>"}``; the signature is HMAC-SHA256 over
``f"{json.dumps(payload)}:{timestamp}:{message_id}"`` with the shared secret (:75-84); a message
older than the TTL (300 s, env DART_MSG_TTL) or with a wrong signature is refused (:150-160);
lists of numbers come back as ndarrays (:209-224).  Bytes produced here verify under the
reference's `SecureSerializer.deserialize` and vice versa (tests/golden/wire.npz was written by
the reference class).  Host-side Python: this is I/O, not the solve.
"""
from __future__ import annotations

import hashlib
import hmac
import json
import os
import time
from typing import Any, Dict, List, Optional

import numpy as np


class WireError(ValueError):
    """Malformed, expired or wrongly signed message (the reference raises CommunicationError)."""


def _plain(obj: Any) -> Any:
    """ndarrays / NumPy scalars -> lists / Python numbers, recursively (:101-105, :176-207)."""
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, np.integer):
        return int(obj)
    if isinstance(obj, np.floating):
        return float(obj)
    if isinstance(obj, dict):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    return obj


def _restore(obj: Any, depth: int = 0, max_depth: int = 100) -> Any:
    """A list made only of numbers becomes an ndarray; everything else keeps its shape (:209-224)."""
    if depth > max_depth:
        raise WireError(f"Maximum recursion depth {max_depth} exceeded during deserialization")
    if isinstance(obj, list):
        if all(isinstance(x, (int, float)) for x in obj):
            return np.array(obj)
        return [_restore(x, depth + 1, max_depth) for x in obj]
    if isinstance(obj, dict):
        return {k: _restore(v, depth + 1, max_depth) for k, v in obj.items()}
    return obj


class SignedEnvelope:
    """Same constructor rules as the reference serializer (:39-68): explicit key, else
    DART_ZMQ_SECRET, else (test mode only) a random one; TTL from the argument, DART_MSG_TTL or 300 s."""

    def __init__(self, secret_key: Optional[str] = None, test_mode: bool = False,
                 message_ttl: Optional[int] = None):
        env_secret = os.getenv("DART_ZMQ_SECRET")
        env_mode = os.getenv("DART_ENVIRONMENT", "development")
        self._test_mode = test_mode or env_mode in ("test", "testing")
        if secret_key:
            self.secret_key = secret_key
        elif env_secret:
            self.secret_key = env_secret
        elif not self._test_mode:
            raise WireError("DART_ZMQ_SECRET must be set in non-test environments for secure ZMQ communication.")
        else:
            import secrets
            self.secret_key = secrets.token_urlsafe(32)
        self._counter = 0
        if message_ttl is not None:
            self._ttl = message_ttl
        else:
            try:
                self._ttl = int(os.getenv("DART_MSG_TTL") or 300)
            except ValueError:
                self._ttl = 300

    def _sign(self, data_json: str, timestamp: float, message_id: str) -> str:
        msg = f"{data_json}:{timestamp}:{message_id}"
        return hmac.new(self.secret_key.encode("utf-8"), msg.encode("utf-8"), hashlib.sha256).hexdigest()

    def serialize(self, obj: Any, *, timestamp: Optional[float] = None, message_id: Optional[str] = None) -> bytes:
        """-> signed message bytes.  `timestamp` / `message_id` default to now and the next
        ``msg_<counter>_<pid>`` (:70-73); they are arguments so that a recorded message can be
        reproduced byte for byte."""
        payload = _plain(obj)
        if timestamp is None:
            timestamp = time.time()
        if message_id is None:
            self._counter += 1
            message_id = f"msg_{self._counter}_{os.getpid()}"
        signature = self._sign(json.dumps(payload), timestamp, message_id)
        return json.dumps({"data": payload, "signature": signature, "timestamp": timestamp,
                           "message_id": message_id}).encode("utf-8")

    def deserialize(self, data: bytes, *, now: Optional[float] = None) -> Any:
        try:
            msg = json.loads(data.decode("utf-8"))
            if not isinstance(msg, dict) or set(msg) != {"data", "signature", "timestamp", "message_id"}:
                raise TypeError("not a signed message")
        except (json.JSONDecodeError, TypeError, UnicodeDecodeError) as e:
            raise WireError(f"Invalid message format: {e}")
        if (time.time() if now is None else now) - msg["timestamp"] > self._ttl:
            raise WireError("Message too old")
        expected = self._sign(json.dumps(msg["data"]), msg["timestamp"], msg["message_id"])
        if not isinstance(msg["signature"], str) or not hmac.compare_digest(msg["signature"], expected):
            raise WireError("Message signature verification failed")
        return _restore(msg["data"])


# ---- payloads -------------------------------------------------------------------------------
def trajectory_payload(traj) -> Dict[str, Any]:
    """The cloud node's answer to a trajectory request (main_improved_threelayer.py:116-124)."""
    return {"positions": np.asarray(traj.positions).tolist(),
            "velocities": None if traj.velocities is None else np.asarray(traj.velocities).tolist(),
            "timestamps": np.asarray(traj.timestamps).tolist()}


def batch_trajectory_payloads(sol, t0, dt: float, ids: Optional[List[Any]] = None) -> List[Dict[str, Any]]:
    """One trajectory payload per problem of a batched solve (HostSolution / BatchSolution.numpy()):
    timestamps t0 + k dt as `_create_trajectory_from_solution` builds them (se3_mpc_planner.py:656-675).
    `t0`: a scalar or one start time per problem; `ids` adds a "drone_id" to each payload."""
    P, V = np.asarray(sol.positions), np.asarray(sol.velocities)
    B, N = P.shape[0], P.shape[1]
    t0 = np.broadcast_to(np.asarray(t0, dtype=np.float64), (B,))
    steps = np.arange(N) * dt
    out = []
    for b in range(B):
        d = {"positions": P[b].tolist(), "velocities": V[b].tolist(), "timestamps": (t0[b] + steps).tolist()}
        if ids is not None:
            d["drone_id"] = ids[b]
        out.append(d)
    return out


def trajectory_from_payload(d: Dict[str, Any]):
    """Inverse of `trajectory_payload` on a *deserialized* payload -> types.Trajectory."""
    from .types import Trajectory
    vel = d.get("velocities")
    return Trajectory(timestamps=np.asarray(d["timestamps"], dtype=np.float64),
                      positions=np.asarray(d["positions"], dtype=np.float64).reshape(-1, 3),
                      velocities=None if vel is None else np.asarray(vel, dtype=np.float64).reshape(-1, 3))
