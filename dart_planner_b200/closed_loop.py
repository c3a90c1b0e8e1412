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

    def _launch(self, lo: int, hi: int, stream, track_counters: bool):
        """One replanning step of drones [lo, hi) (one launch): solve, store, advance the plant."""
        torch = _torch()
        es = 8 * self.ld
        base = self.state.data_ptr() + 8 * lo
        meta = self.meta.data_ptr() + 4 * lo
        with torch.cuda.device(self.device):
            rc = self._lib.dart_se3mpc_closed_loop_step(
                C.byref(self.params), hi - lo, self.ld, base, base + 3 * es, base + 6 * es, None,
                self.x.data_ptr() + 8 * lo,
                0 if self.steps_done == 0 else (2 if self.lateral_thrust_is_zero else 1), self.cost.data_ptr() + 8 * lo,
                meta, meta + 4 * self.ld, meta + 8 * self.ld,
                self.plant_dt, stream.cuda_stream)
        _cabi.check(rc, "dart_se3mpc_closed_loop_step")
        if track_counters:
            with torch.cuda.stream(stream):
                self.nfev_total[lo:hi] += self.meta[1, lo:hi]

    def step(self, stream=None, track_counters: bool = True):
        """One replanning step of the whole population (one launch)."""
        torch = _torch()
        self._launch(0, self.B, stream or torch.cuda.current_stream(self.device), track_counters)
        self.steps_done += 1

    def default_parts(self) -> int:
        """Sub-populations `run` keeps in flight (measured, tools/closed_loop_parts.py, 100 replans on
        one B200, ms for 1 / 2 / 3 / 4 parts): 65 536 drones 22.0 / 21.6 / 21.6 / 21.7; 32 768: 11.5 /
        10.8 / 10.8 / 11.1; 16 384: 6.8 / 5.6 / 5.5 / 5.6; 8 192 (one GPU's share of 65 536 on eight):
        4.6 / 3.5 / 3.3 / 3.6 -- the fewer waves of blocks a step has, the more its idle tail costs."""
        return 2 if self.B >= 49152 else (3 if self.B >= 6144 else 1)

    def run(self, steps: int, stream=None, track_counters: bool = True, record: bool = False,
            parts: Optional[int] = None):
        """`steps` replans.  record=True returns the (steps, B, 3) position history (host).

        A drone's step k+1 depends on ITS step k only, so the population is replanned as `parts`
        independent sub-populations (contiguous index ranges, a CUDA stream each): while the last,
        partly filled wave of one sub-population's step drains, the next step of another one
        already fills the free SMs -- with one launch per step for everybody, every step ends
        with that idle tail (65 536 drones: 9.2 waves of blocks per step; 8 192 per GPU on eight
        GPUs: 1.15).  Same launches, same results per drone; only their order on the machine
        changes.  The library is told how many problems are in flight so that the sub-population
        launches are served by the throughput build like the whole population would be."""
        torch = _torch()
        stream = stream or torch.cuda.current_stream(self.device)
        parts = self.default_parts() if parts is None else max(1, int(parts))
        if record or parts == 1 or steps <= 0:
            hist = []
            for _ in range(steps):
                self.step(stream, track_counters)
                if record:
                    hist.append(self.positions().cpu().numpy())
            return np.stack(hist) if record else None
        # contiguous ranges, multiples of 64 drones (whole 512-byte row segments)
        per = (self.B // parts + 63) // 64 * 64
        ranges = [(a, min(a + per, self.B)) for a in range(0, self.B, per)]
        if getattr(self, "_side", None) is None or len(self._side) < len(ranges) - 1:
            self._side = [torch.cuda.Stream(self.device) for _ in range(len(ranges) - 1)]
        streams = [stream] + self._side[: len(ranges) - 1]
        start = torch.cuda.Event()
        start.record(stream)
        for st in streams[1:]:
            st.wait_event(start)
        self._lib.dart_se3mpc_set_inflight_hint(self.B)
        try:
            for _ in range(steps):
                for (lo, hi), st in zip(ranges, streams):
                    self._launch(lo, hi, st, track_counters)
                self.steps_done += 1
        finally:
            self._lib.dart_se3mpc_set_inflight_hint(0)
        for st in streams[1:]:
            done = torch.cuda.Event()
            done.record(st)
            stream.wait_event(done)
        return None

    def positions(self):
        return self.state[0:3, : self.B].t()

    def velocities(self):
        return self.state[3:6, : self.B].t()

    def goals(self):
        return self.state[6:9, : self.B].t()

    def solution(self):
        """(B, 9N) current solutions in the reference's packed order."""
        return self.x[:, : self.B].t()
