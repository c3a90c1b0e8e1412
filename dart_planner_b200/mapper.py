"""Dense device occupancy grid with the query surface of the reference's
``ExplicitGeometricMapper`` (perception/explicit_geometric_mapper.py), batched.

The reference keeps a sparse dict keyed by ``floor(p / resolution)`` with prior 0.5 for
unknown voxels (:154-169).  Here the map is a dense fp32 grid [nz][ny][nx] resident in HBM
(256^3 = 64 MiB, L2-resident on B200); keys outside the grid read the prior, as a dict miss.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _cabi


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("dart_planner_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch


class DenseOccupancyGrid:
    def __init__(self, shape: Tuple[int, int, int] = (256, 256, 256),
                 origin_voxel: Tuple[int, int, int] = (-128, -128, -128),
                 resolution: float = 0.2, prior: float = 0.5, device=None, max_range: float = 50.0):
        torch = _torch()
        self.nx, self.ny, self.nz = (int(s) for s in shape)
        self.origin_voxel = tuple(int(o) for o in origin_voxel)
        self.resolution = float(resolution)
        self.prob_prior = float(prior)
        self.prob_hit, self.prob_miss = 0.7, 0.4        # explicit_geometric_mapper.py:80-81
        self.max_range = float(max_range)               # :66
        self.total_observations = 0
        self._counts = None                              # update_map scratch (uint64 per cell)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.occ = torch.full((self.nz, self.ny, self.nx), prior, dtype=torch.float32, device=self.device)

    # -- plumbing ------------------------------------------------------------------------
    def _grid(self) -> _cabi.Grid:
        return _cabi.Grid(self.nx, self.ny, self.nz, *self.origin_voxel, self.resolution,
                          self.prob_prior, self.occ.data_ptr())

    def _stream(self, stream):
        torch = _torch()
        return (stream or torch.cuda.current_stream(self.device)).cuda_stream

    def _soa(self, a, rows):
        """(B, rows) array-like -> contiguous (rows, B) float64 device tensor."""
        torch = _torch()
        t = torch.as_tensor(np.asarray(a, np.float64) if not torch.is_tensor(a) else a,
                            dtype=torch.float64).to(self.device)
        return t.reshape(-1, rows).t().contiguous()

    # -- reference surface -----------------------------------------------------------------
    def world_to_voxel(self, position) -> Tuple[int, int, int]:
        return tuple(int(v) for v in np.floor(np.asarray(position, float) / self.resolution).astype(int))

    def add_obstacle(self, center, radius: float, value: float = 0.9) -> None:
        self.add_obstacles(np.asarray(center, float).reshape(1, 3), [radius], value)

    def add_obstacles(self, centers, radii, value: float = 0.9, stream=None) -> None:
        """Rasterise spheres exactly like add_obstacle (:399-423)."""
        torch = _torch()
        c = self._soa(centers, 3)
        r = torch.as_tensor(np.asarray(radii, np.float64), dtype=torch.float64).to(self.device).contiguous()
        g = self._grid()
        rc = _cabi.lib().dart_map_add_spheres(C.byref(g), self.occ.data_ptr(), int(r.numel()),
                                              c.data_ptr(), r.data_ptr(), float(value), self._stream(stream))
        _cabi.check(rc, "dart_map_add_spheres")

    def query_occupancy_batch(self, positions, stream=None):
        """positions (B,3) -> (B,) float64 device tensor (query_occupancy_batch :171-182)."""
        torch = _torch()
        pos = self._soa(positions, 3)
        B = pos.shape[1]
        out = torch.empty(B, dtype=torch.float64, device=self.device)
        g = self._grid()
        rc = _cabi.lib().dart_map_query_batch(C.byref(g), B, B, pos.data_ptr(), out.data_ptr(),
                                              self._stream(stream))
        _cabi.check(rc, "dart_map_query_batch")
        return out

    def query_occupancy(self, position) -> float:
        return float(self.query_occupancy_batch(np.asarray(position, float).reshape(1, 3))[0])

    def is_collision(self, position, threshold: float = 0.6) -> bool:
        return self.query_occupancy(position) > threshold

    def trajectories_safe_soa(self, positions_soa, B: int, npos: int, safety_margin: float = 1.0,
                              threshold: float = 0.6, out=None, stream=None):
        """positions_soa: (3*npos, ld) device tensor with rows 3k+c (e.g. the first 3N rows of a
        BatchSolution.out) -> (ld,) int32 first-collision index, -1 = safe."""
        torch = _torch()
        ld = positions_soa.shape[1]
        assert positions_soa.is_cuda and positions_soa.dtype == torch.float64 and positions_soa.is_contiguous()
        assert positions_soa.shape[0] >= 3 * npos
        if out is None:
            out = torch.empty(ld, dtype=torch.int32, device=self.device)
        g = self._grid()
        rc = _cabi.lib().dart_map_traj_safe_batch(C.byref(g), B, ld, npos, positions_soa.data_ptr(),
                                                  float(safety_margin), float(threshold),
                                                  out.data_ptr(), self._stream(stream))
        _cabi.check(rc, "dart_map_traj_safe_batch")
        return out

    def are_trajectories_safe(self, positions, safety_margin: float = 1.0, threshold: float = 0.6):
        """positions (B, npos, 3) -> (B,) int32 first-collision index, -1 = safe."""
        torch = _torch()
        t = torch.as_tensor(np.asarray(positions, np.float64) if not torch.is_tensor(positions) else positions,
                            dtype=torch.float64).to(self.device)
        B, npos, _ = t.shape
        soa = t.reshape(B, 3 * npos).t().contiguous()
        return self.trajectories_safe_soa(soa, B, npos, safety_margin, threshold)[:B]

    def is_trajectory_safe(self, positions, safety_margin: float = 1.0, threshold: float = 0.6):
        """Reference signature (:195-219): returns (is_safe, first_collision_index)."""
        idx = int(self.are_trajectories_safe(np.asarray(positions, float)[None], safety_margin, threshold)[0])
        return idx < 0, idx

    def update_map(self, positions, directions, hit_distances, max_ranges=50.0, stream=None):
        """Batched `update_map` (:100-152) for one scan: B observations given as arrays
        (SensorObservation.position / direction / hit_distance / max_range; hit_distance None or
        NaN = no return).  Two launches: ray walk with per-voxel visit counters, then the Bayes
        rule applied per voxel.  Returns the reference's counters dict."""
        torch = _torch()
        start = self._soa(positions, 3)
        d = self._soa(directions, 3)
        B = start.shape[1]
        if torch.is_tensor(hit_distances):
            hit_t = hit_distances.to(self.device, torch.float64).reshape(-1)
        else:
            hd = np.asarray(hit_distances)
            if hd.dtype == object:      # the reference's Optional[float] per observation: None = no return
                hd = np.array([np.nan if h is None else float(h) for h in hd.reshape(-1)])
            hit_t = torch.as_tensor(np.ascontiguousarray(hd, np.float64).reshape(-1)).to(self.device)
        mr = torch.as_tensor(np.broadcast_to(np.asarray(max_ranges, np.float64), (B,)).copy()).to(self.device)
        if self._counts is None:
            self._counts = torch.zeros(self.nx * self.ny * self.nz, dtype=torch.int64, device=self.device)
        updated = torch.zeros(1, dtype=torch.int64, device=self.device)
        g = self._grid()
        rc = _cabi.lib().dart_map_update_batch(
            C.byref(g), self.occ.data_ptr(), self._counts.data_ptr(), B, B, start.data_ptr(), d.data_ptr(),
            hit_t.data_ptr(), mr.data_ptr(), self.max_range, self.prob_hit, self.prob_miss,
            updated.data_ptr(), self._stream(stream))
        _cabi.check(rc, "dart_map_update_batch")
        self.total_observations += B
        return {"updated_voxels": int(updated.item()), "observations_processed": B,
                "total_voxels": int((self.occ != self.prob_prior).sum().item())}

    def trace_rays(self, starts, directions, distances, max_vox: int = 0, stream=None):
        """_trace_ray (:250-309) for B rays.  Returns (count (B,) int32, voxels (max_vox,3,B) int32
        or None): voxels[:count[b], :, b] is the reference's list for ray b."""
        torch = _torch()
        s, d = self._soa(starts, 3), self._soa(directions, 3)
        B = s.shape[1]
        dist = torch.as_tensor(np.asarray(distances, np.float64) if not torch.is_tensor(distances) else distances,
                               dtype=torch.float64).to(self.device).reshape(B).contiguous()
        count = torch.empty(B, dtype=torch.int32, device=self.device)
        vox = torch.zeros((max_vox, 3, B), dtype=torch.int32, device=self.device) if max_vox > 0 else None
        rc = _cabi.lib().dart_map_trace_ray_batch(self.resolution, B, B, s.data_ptr(), d.data_ptr(),
                                                  dist.data_ptr(), int(max_vox), count.data_ptr(),
                                                  None if vox is None else vox.data_ptr(),
                                                  self._stream(stream))
        _cabi.check(rc, "dart_map_trace_ray_batch")
        return count, vox

    def _trace_ray(self, start, direction, distance: float, max_vox: int = 4096):
        """Reference signature: list of voxel-index tuples for one ray."""
        count, vox = self.trace_rays(np.asarray(start, float)[None], np.asarray(direction, float)[None],
                                     [distance], max_vox=max_vox)
        n = int(count[0])
        if n > max_vox:
            raise ValueError(f"ray visits {n} voxels > max_vox={max_vox}")
        v = vox[:n, :, 0].cpu().numpy()
        return [tuple(int(c) for c in row) for row in v]
