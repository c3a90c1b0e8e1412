"""Host-side derivation of the arrays the reference computes from an optimised thrust sequence
(`_extract_solution_from_result` / `_compute_attitudes_and_rates`,
src/dart_planner/planning/se3_mpc_planner.py:582-654), vectorised over a batch.

Used by ``HostSolution`` for *solution rows* (``DART_ROWS_SOLUTION``): the kernel returns what
``scipy.optimize.minimize`` returns (x, cost, counters: half the PCIe bytes of a full row) and
these pure functions of the thrust rows of x are evaluated on first access.  The solve itself
never runs here.
"""
from __future__ import annotations

import numpy as np


def derive_from_thrust(T: np.ndarray, dt: float, mass: float, gravity: float):
    """T: (B, N, 3) thrust vectors -> (accelerations (B,N,3), attitudes (B,N,3) roll/pitch/yaw,
    body_rates (B,N,3), thrusts (B,N)).

    Per step (:615-652): a step with |T| <= 1e-6 gets zero attitude and rates and does NOT
    advance ``prev_R``; b1 = (1,0,0) x b3 falls back to (1,0,0) when its norm is <= 1e-6; the
    body rate is vee(R^T (R - prev_R) / dt) against the previous VALID step's rotation."""
    T = np.asarray(T, dtype=np.float64)
    B, N, _ = T.shape
    acc = T / mass - np.array([0.0, 0.0, gravity])
    thr = np.sqrt((T * T).sum(axis=2))
    valid = thr > 1e-6                     # NaN magnitudes are not valid (comparison is False)
    with np.errstate(all="ignore"):
        b3 = T / thr[..., None]
        # yaw_vector = (1, 0, 0):  b1 = yaw_vector x b3 = (0*b3z - 0*b3y, 0*b3x - 1*b3z, 1*b3y - 0*b3x)
        b1 = np.stack([0.0 * b3[..., 2] - 0.0 * b3[..., 1], 0.0 * b3[..., 0] - b3[..., 2],
                       b3[..., 1] - 0.0 * b3[..., 0]], axis=-1)
        n1 = np.sqrt((b1 * b1).sum(axis=2))
        ok1 = n1 > 1e-6
        b1 = np.where(ok1[..., None], b1 / n1[..., None], np.array([1.0, 0.0, 0.0]))
        b2 = np.cross(b3, b1)
        R = np.stack([b1, b2, b3], axis=-1)           # columns b1 b2 b3 -> R[b, k, row, col]
        att = np.stack([np.arctan2(R[..., 2, 1], R[..., 2, 2]), np.arcsin(-R[..., 2, 0]),
                        np.arctan2(R[..., 1, 0], R[..., 0, 0])], axis=-1)
    att = np.where(valid[..., None], att, 0.0)
    rates = np.zeros((B, N, 3))
    prev = np.zeros((B, 3, 3))
    have = np.zeros(B, dtype=bool)
    for i in range(N):
        Ri, vi = R[:, i], valid[:, i]
        with np.errstate(all="ignore"):
            Rd = (Ri - prev) / dt
            om = np.einsum("bji,bjk->bik", Ri, Rd)    # R^T Rdot
        w = np.stack([om[:, 2, 1], om[:, 0, 2], om[:, 1, 0]], axis=-1)
        rates[:, i] = np.where((vi & have)[:, None], w, 0.0)
        prev = np.where(vi[:, None, None], Ri, prev)
        have |= vi
    return acc, att, rates, thr
