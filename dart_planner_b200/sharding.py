"""Sharding of a batch of independent SE(3)-MPC problems across the GPUs of one box.

Every problem is independent and the map is read-only, so the batch is cut
into contiguous slices of the problem index, one slice per rank (one process per GPU), each
rank solves its slice with a replica of the parameters / map, and the results return through
ONE collective at the end: a gather of the fp64 result block and the int32 counter block to
rank 0 over NCCL (NVLink 5 / NVSwitch).  There is no collective inside the solve.

The partition and gather plumbing is backend-agnostic (`gloo` in the CPU tests, `nccl` on the
box); the solve itself is always the CUDA path -- ``solve_fn`` exists so the CPU tests can
exercise the plumbing with a stand-in, the default raises without a CUDA device.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np


def shard_range(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced slice [lo, hi) of the problem index owned by `rank`.
    The first ``B % world`` ranks hold one extra problem; empty slices are allowed."""
    if world < 1 or not 0 <= rank < world or B < 0:
        raise ValueError(f"bad shard request B={B} world={world} rank={rank}")
    q, r = divmod(B, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_counts(B: int, world: int) -> List[int]:
    return [shard_range(B, world, r)[1] - shard_range(B, world, r)[0] for r in range(world)]


def gather_columns(local, counts: List[int], dst: int = 0, group=None):
    """Gather column blocks of a batch-major SoA tensor: every rank holds ``(rows, counts[rank])``;
    rank `dst` receives ``(rows, sum(counts))`` with the slices in rank order, others ``None``.
    One collective: blocks are padded to the widest slice so a plain ``gather`` serves ragged
    partitions (NCCL has no gatherv)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(counts) == world and local.shape[1] == counts[rank]
    rows, width = local.shape[0], max(max(counts), 1)
    send = local
    if local.shape[1] != width:
        send = torch.zeros((rows, width), dtype=local.dtype, device=local.device)
        send[:, : local.shape[1]] = local
    send = send.contiguous()
    if world == 1:
        return local
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:, :c] for b, c in zip(bufs, counts)], dim=1)


class ShardedSolver:
    """Solve a global batch sharded by problem index; results land on rank `dst`.

    >>> solver = ShardedSolver(params)          # after dist.init_process_group("nccl")
    >>> sol = solver.solve(p0, v0, goal)         # every rank passes the same global arrays (or
    ...                                          # only its slice with ``presliced=True``)
    ``sol`` is a HostSolution on rank `dst`, ``None`` elsewhere.
    """

    def __init__(self, params, *, dst: int = 0, group=None,
                 solve_fn: Optional[Callable] = None, outputs: str = "all"):
        self.params = params
        self.dst = dst
        self.group = group
        self.outputs = outputs
        self._solve_fn = solve_fn
        self._ws = None

    def _solve_local(self, p0, v0, goal):
        """-> (out (19N+1, b) float64, meta (4, b) int32) torch tensors for the local slice."""
        if self._solve_fn is not None:
            return self._solve_fn(self.params, p0, v0, goal)
        from .planner import BatchWorkspace  # raises without CUDA: no CPU fallback

        b = len(p0)
        if self._ws is None or self._ws.B != b:
            self._ws = BatchWorkspace(self.params, max(b, 1), pinned=False, outputs=self.outputs)
        if b == 0:
            return self._ws.out[:, :0], self._ws.meta[:, :0]
        self._ws.set_inputs_device(p0, v0, goal)
        sol = self._ws.solve_device()
        return sol.out[:, :b], sol.meta[:, :b]

    def solve(self, p0, v0, goal, presliced: bool = False, global_B: Optional[int] = None):
        import torch.distributed as dist

        from .planner import HostSolution

        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        p0, v0, goal = (np.asarray(a, np.float64).reshape(-1, 3) for a in (p0, v0, goal))
        if presliced:
            if global_B is None:
                raise ValueError("presliced=True needs global_B")
            B = int(global_B)
            lo, hi = shard_range(B, world, rank)
            if len(p0) != hi - lo:
                raise ValueError(f"rank {rank} owns {hi - lo} problems, got {len(p0)}")
        else:
            B = len(p0)
            lo, hi = shard_range(B, world, rank)
            p0, v0, goal = p0[lo:hi], v0[lo:hi], goal[lo:hi]
        out, meta = self._solve_local(p0, v0, goal)
        counts = shard_counts(B, world)
        if world > 1:
            out = gather_columns(out, counts, self.dst, self.group)
            meta = gather_columns(meta, counts, self.dst, self.group)
        if rank != self.dst:
            return None
        N = int(self.params.horizon)
        if out.is_cuda:      # transpose on the device, one contiguous read-back
            return HostSolution.from_rows(N, out.t().contiguous().cpu().numpy(), meta.cpu().numpy())
        return HostSolution.from_blocks(N, out.cpu().numpy(), meta.cpu().numpy())
