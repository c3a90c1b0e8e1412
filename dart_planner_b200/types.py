"""`DroneState` / `Trajectory` with the reference's field names (common/types.py:63-139).

Fields hold plain SI ndarrays (or anything exposing ``.magnitude``, which is stripped): the
reference's solve is only executable on SI magnitudes.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .config import to_si


def _vec3(v, unit):
    a = np.asarray(to_si(v, unit), dtype=np.float64).reshape(-1)
    if a.shape != (3,):
        raise ValueError(f"expected a 3-vector, got shape {a.shape}")
    return a


@dataclass
class DroneState:
    timestamp: float
    position: np.ndarray = field(default_factory=lambda: np.zeros(3))
    velocity: np.ndarray = field(default_factory=lambda: np.zeros(3))
    attitude: np.ndarray = field(default_factory=lambda: np.zeros(3))
    angular_velocity: np.ndarray = field(default_factory=lambda: np.zeros(3))
    motor_rpms: Optional[np.ndarray] = field(default_factory=lambda: np.zeros(4))

    def __post_init__(self):
        self.position = _vec3(self.position, "m")
        self.velocity = _vec3(self.velocity, "m/s")
        self.attitude = _vec3(self.attitude, "rad")
        self.angular_velocity = _vec3(self.angular_velocity, "rad/s")


@dataclass
class Trajectory:
    timestamps: np.ndarray
    positions: np.ndarray
    velocities: Optional[np.ndarray] = None
    accelerations: Optional[np.ndarray] = None
    attitudes: Optional[np.ndarray] = None   # roll, pitch, yaw (rad)
    body_rates: Optional[np.ndarray] = None  # rad/s
    thrusts: Optional[np.ndarray] = None     # N
    yaws: Optional[np.ndarray] = None
    yaw_rates: Optional[np.ndarray] = None
