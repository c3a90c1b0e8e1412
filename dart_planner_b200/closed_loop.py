"""Closed-loop receding-horizon simulation on resident device state (BASELINE configs[4]).

Every drone replans at a fixed rate: solve from the current (p, v) with the warm start of
se3_mpc_planner.py:294-327 (the reference builds it but never stores `last_solution`; here the
previous solution is kept), then the state advances with the planner's own
model (:430-431, :445-459) driven by the first control of the new solution.  One kernel launch
per replanning step for the whole population (`dart_se3mpc_closed_loop_step`): state, goals and
solutions never leave HBM; only what the caller asks for is copied out.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _cabi


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("dart_planner_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch


class ClosedLoopSim:
    """B drones, fixed goals, replanning every `plant_dt` seconds.

    >>> sim = ClosedLoopSim(params, B, plant_dt=0.1)
    >>> sim.reset(p0, v0, goals)
    >>> sim.run(100)                  # 100 launches, state stays on the device
    >>> sim.positions(), sim.velocities(), sim.nfev_total
    """

    def __init__(self, params: _cabi.Params, B: int, plant_dt: Optional[float] = None, device=None):
        torch = _torch()
        self.params = params
        self.N = int(params.horizon)
        self.B = int(B)
        self.ld = max(32, (self.B + 31) // 32 * 32)
        self.plant_dt = float(params.dt if plant_dt is None else plant_dt)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        f8 = dict(dtype=torch.float64, device=self.device)
        self.state = torch.zeros((9, self.ld), **f8)            # rows p xyz | v xyz | goal xyz
        self.x = torch.zeros((9 * self.N, self.ld), **f8)       # current solution, reference order
        self.cost = torch.zeros(self.ld, **f8)
        self.meta = torch.zeros((3, self.ld), dtype=torch.int32, device=self.device)  # nit, nfev, status
        self.nfev_total = torch.zeros(self.ld, dtype=torch.int64, device=self.device)
        self.steps_done = 0
        # the solutions held in self.x come from cold starts of this solver, so their lateral
        # thrust is exactly zero and stays zero: warm starts may use the 7-slot kernel (warm=2;
        # the kernel verifies).  Anyone writing self.x by hand must clear this flag.
        self.lateral_thrust_is_zero = True
        self._lib = _cabi.lib()

    def reset(self, p0, v0, goals):
        torch = _torch()
        for i, a in enumerate((p0, v0, goals)):
            t = torch.as_tensor(np.asarray(a, np.float64) if not torch.is_tensor(a) else a,
                                dtype=torch.float64).to(self.device)
            self.state[3 * i: 3 * i + 3, : self.B] = t.reshape(self.B, 3).t()
        self.x.zero_()
        self.nfev_total.zero_()
        self.steps_done = 0
        self.lateral_thrust_is_zero = True

    def step(self, stream=None, track_counters: bool = True):
        """One replanning step (one launch): solve, store the solution, advance the plant."""
        torch = _torch()
        stream = stream or torch.cuda.current_stream(self.device)
        es = 8 * self.ld
        base = self.state.data_ptr()
        with torch.cuda.device(self.device):
            rc = self._lib.dart_se3mpc_closed_loop_step(
                C.byref(self.params), self.B, self.ld, base, base + 3 * es, base + 6 * es, None,
                self.x.data_ptr(),
                0 if self.steps_done == 0 else (2 if self.lateral_thrust_is_zero else 1), self.cost.data_ptr(),
                self.meta.data_ptr(), self.meta.data_ptr() + 4 * self.ld, self.meta.data_ptr() + 8 * self.ld,
                self.plant_dt, stream.cuda_stream)
        _cabi.check(rc, "dart_se3mpc_closed_loop_step")
        if track_counters:
            with torch.cuda.stream(stream):
                self.nfev_total += self.meta[1]
        self.steps_done += 1

    def run(self, steps: int, stream=None, track_counters: bool = True, record: bool = False):
        """`steps` replans.  record=True returns the (steps, B, 3) position history (host)."""
        hist = []
        for _ in range(steps):
            self.step(stream, track_counters)
            if record:
                hist.append(self.positions().cpu().numpy())
        return np.stack(hist) if record else None

    def positions(self):
        return self.state[0:3, : self.B].t()

    def velocities(self):
        return self.state[3:6, : self.B].t()

    def goals(self):
        return self.state[6:9, : self.B].t()

    def solution(self):
        """(B, 9N) current solutions in the reference's packed order."""
        return self.x[:, : self.B].t()
