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


def gather_row_blocks(local, counts: List[int], dst: int = 0, group=None, recv=None):
    """Gather problem-major row blocks: every rank holds ``(counts[rank], stride)`` (one contiguous
    result row per problem); rank `dst` receives them in rank order.  ONE collective, written by
    NCCL straight into slices of one receive block (no concatenation): returns ``(block, width)``
    on `dst` -- rank r's rows are ``block[r * width : r * width + counts[r]]`` -- and ``(None,
    width)`` elsewhere.  Slices are padded to the widest one (NCCL has no gatherv)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(counts) == world and local.shape[0] == counts[rank]
    width, stride = max(max(counts), 1), local.shape[1]
    send = local
    if local.shape[0] != width:
        send = torch.zeros((width, stride), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    send = send.contiguous()
    if rank != dst:
        dist.gather(send, None, dst=dst, group=group)
        return None, width
    if recv is None or tuple(recv.shape) != (world * width, stride) or recv.device != local.device:
        recv = torch.empty((world * width, stride), dtype=local.dtype, device=local.device)
    dist.gather(send, [recv[r * width:(r + 1) * width] for r in range(world)], dst=dst, group=group)
    return recv, width


def replicate_map(grid, src: int = 0, group=None):
    """The replicated map of BASELINE configs[3]: every rank holds a `DenseOccupancyGrid` of the same
    geometry on its own GPU (each rank constructs its own), and the cells of rank `src` -- the map
    the mapper built there -- are copied into all of them by ONE broadcast (256^3 float32 cells:
    64 MiB over NVLink).  Read-only afterwards: every rank's solves query their local replica, no
    map traffic crosses a link during a solve.  Returns `grid`; a no-op for a single process."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grid
    mine = (grid.nx, grid.ny, grid.nz, tuple(grid.origin_voxel), grid.resolution, grid.prob_prior, grid.dtype)
    geo = [mine if dist.get_rank(group) == src else None]
    dist.broadcast_object_list(geo, src=src, group=group)
    same = [None] * dist.get_world_size(group)
    dist.all_gather_object(same, tuple(geo[0]) == mine, group=group)      # fail together, not in the broadcast
    if not all(same):
        raise ValueError(f"replicate_map: ranks {[r for r, o in enumerate(same) if not o]} hold a grid of another "
                         f"geometry than rank {src}'s {geo[0]}")
    dist.broadcast(grid.occ, src=src, group=group)
    return grid


class SharedHostBlock:
    """One (B, stride) float64 block in POSIX shared memory, mapped by every rank of the group
    (one process per GPU on one box) and page-locked for each rank's GPU: rank r's kernel writes the
    result rows of ITS slice straight into rows [lo_r, hi_r) over its own PCIe link, and rank `dst`
    reads the whole block as host memory.  No collective moves a result; the N links work in
    parallel instead of everything funnelling through one GPU's link.

    ``register=False`` (CPU plumbing tests) maps the block without page-locking it."""

    def __init__(self, B: int, stride: int, dst: int = 0, group=None, register: bool = True):
        import torch.distributed as dist
        from multiprocessing import shared_memory

        multi = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if multi else 1
        self.rank = dist.get_rank(group) if multi else 0
        self.B, self.stride, self.dst = int(B), int(stride), dst
        self.nbytes = max(self.B, 1) * self.stride * 8
        self.owner = self.rank == dst
        name = [None]
        err = None
        self._shm = None
        self._registered = False
        self.rows = None
        if self.owner:
            try:
                self._shm = shared_memory.SharedMemory(create=True, size=self.nbytes)
                name[0] = self._shm.name
            except Exception as e:      # noqa: BLE001  (reported to every rank below)
                err = e
        if self.world > 1:
            dist.broadcast_object_list(name, src=dst, group=group)
        if err is None and name[0] is None:
            err = RuntimeError("the owning rank could not create the shared result block")
        if err is None:
            try:
                if not self.owner:
                    self._shm = shared_memory.SharedMemory(name=name[0])
                    try:    # the creating rank unlinks the segment; attaching ranks must not (Python < 3.13 tracks them too)
                        from multiprocessing import resource_tracker
                        resource_tracker.unregister(self._shm._name, "shared_memory")
                    except Exception:
                        pass
                self.rows = np.ndarray((max(self.B, 1), self.stride), dtype=np.float64, buffer=self._shm.buf)
                self.ptr = self.rows.ctypes.data
                if register:
                    import torch
                    rc = torch.cuda.cudart().cudaHostRegister(self.ptr, self.nbytes, 1 | 2)   # portable | mapped
                    if int(rc) != 0:
                        raise RuntimeError(f"cudaHostRegister of the shared result block failed ({int(rc)})")
                    self._registered = True
            except Exception as e:      # noqa: BLE001
                err = e
        # Everybody is attached before anybody may finish and unlink -- and a rank that failed must
        # not leave the others waiting in a later collective: the ranks agree on the outcome here
        # and fail (or go on) TOGETHER.
        if self.world > 1:
            oks = [None] * self.world
            dist.all_gather_object(oks, err is None, group=group)
            if err is None and not all(oks):
                err = RuntimeError(f"shared result block: ranks {[r for r, o in enumerate(oks) if not o]} failed to map it")
        if err is not None:
            self.close()
            raise err

    def close(self):
        if getattr(self, "_shm", None) is None:
            return
        if self._registered:
            import torch
            torch.cuda.cudart().cudaHostUnregister(self.ptr)
            self._registered = False
        self.rows = None
        try:
            self._shm.close()
            if self.owner:
                self._shm.unlink()
        except Exception:
            pass
        self._shm = None

    def __del__(self):
        self.close()


class ShardedSolver:
    """Solve a global batch sharded by problem index; results land on rank `dst`.

    >>> solver = ShardedSolver(params)          # after dist.init_process_group("nccl")
    >>> sol = solver.solve(p0, v0, goal)         # every rank passes the same global arrays (or
    ...                                          # only its slice with ``presliced=True``)
    ``sol`` is a HostSolution on rank `dst`, ``None`` elsewhere.

    Data path on the box: each rank's kernel writes one packed result row per problem into device
    memory (``outputs``: "all" 19N+4 doubles, "solution" x | cost | counters, "controls" thrust
    vectors | cost | counters), ONE NCCL gather moves the slices over NVLink into one block on
    `dst`, ONE copy per rank slice brings it into a cached pinned host block, and the HostSolution
    is a set of views of that block (valid until the next ``solve``; ``copy=True`` detaches it).
    ``last_timing`` holds the milliseconds of the stages on `dst`.

    ``transport="host_block"`` replaces the gather: the ranks share ONE page-locked host block
    (`SharedHostBlock`) and every rank's kernel writes its slice of result rows straight into it
    over its own PCIe link (zero-copy, full 128-byte lines); a barrier ends the call and `dst` reads
    host memory.  No result crosses NVLink and nothing funnels through `dst`'s link: 1 Mi problems
    reach rank 0's host memory 3-5x sooner (bench.py, `sharded_configs`).  Needs N <= 25 (rows).

    ``set_map(grid, safety_margin, collision_threshold)`` attaches this rank's replica of the map
    (`replicate_map`): every solve then runs the fused `is_trajectory_safe` check on its solved
    positions against the LOCAL replica and the rows carry its result (`HostSolution.first_hit`)."""

    def __init__(self, params, *, dst: int = 0, group=None,
                 solve_fn: Optional[Callable] = None, outputs: str = "all",
                 rows_fn: Optional[Callable] = None, transport: str = "gather"):
        if transport not in ("gather", "host_block"):
            raise ValueError("transport must be 'gather' or 'host_block'")
        self.params = params
        self.dst = dst
        self.group = group
        self.outputs = outputs
        self.transport = transport
        self._block = None
        self._solve_fn = solve_fn
        self._rows_fn = rows_fn
        self._ws = None
        self._recv = None
        self._pinned = None
        self._map = None
        self.last_timing = {}

    def set_map(self, grid, safety_margin: float = 1.0, collision_threshold: float = 0.6):
        """This rank's replica of the occupancy grid (None detaches it): the solves of the CUDA row
        path also run the fused trajectory safety check against it (explicit_geometric_mapper.py:
        195-219); the result travels in the rows ("all" and "solution" rows; controls rows carry none)."""
        self._map = None if grid is None else (grid, float(safety_margin), float(collision_threshold))
        if self._ws is not None:
            self._ws.set_map(*(self._map or (None,)))

    # -- local solves -----------------------------------------------------------------------
    def _solve_local(self, p0, v0, goal):
        """SoA stand-in path (CPU plumbing tests): (out (19N+1, b), meta (4, b)) torch tensors."""
        return self._solve_fn(self.params, p0, v0, goal)

    def _workspace(self, b):
        from .planner import BatchWorkspace  # raises without CUDA: no CPU fallback

        if self._ws is None or self._ws.B != max(b, 1):
            self._ws = BatchWorkspace(self.params, max(b, 1), pinned=False, outputs=self.outputs)
            if self._map is not None:
                self._ws.set_map(*self._map)
        return self._ws

    def _rows_local(self, p0, v0, goal):
        """-> (b, stride) tensor of packed result rows for the local slice (device tensor on the box)."""
        if self._rows_fn is not None:
            return self._rows_fn(self.params, p0, v0, goal, self.outputs)
        ws = self._workspace(len(p0))
        if len(p0):
            ws.set_inputs_device(p0, v0, goal)
        return self._rows_resident(len(p0))

    def _rows_resident(self, b):
        ws = self._workspace(b)
        rows = ws.solve_rows_device()
        return rows[:b]             # an empty slice still takes part in the gather

    def rows_supported(self) -> bool:
        if self._solve_fn is not None:
            return False
        if self._rows_fn is not None:
            return True
        return self._workspace(1).row_stride > 0

    # -- the call ---------------------------------------------------------------------------
    def _slice(self, p0, v0, goal, presliced, global_B):
        import torch.distributed as dist

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
        return B, world, rank, p0, v0, goal

    def stage(self, p0, v0, goal, presliced: bool = False, global_B: Optional[int] = None):
        """Upload this rank's slice of the inputs (CUDA path); `run()` then solves the resident
        slice and gathers -- the split bench.py uses to time the solve + gather on the device."""
        B, world, rank, p0, v0, goal = self._slice(p0, v0, goal, presliced, global_B)
        if not self.rows_supported() or self._rows_fn is not None:
            raise RuntimeError("stage()/run() is the CUDA row path")
        if len(p0):
            self._workspace(len(p0)).set_inputs_device(p0, v0, goal)
        self._staged = (B, world, rank, len(p0))

    def run(self, copy: bool = False):
        B, world, rank, b = self._staged
        if self.transport == "host_block":
            return self._run_host_block(B, world, rank, b, copy)
        return self._finish(self._rows_resident(b), B, world, rank, copy, None)

    # -- host-block transport ---------------------------------------------------------------
    def _shared_block(self, B, stride, register=True):
        blk = self._block
        if blk is None or blk.B != B or blk.stride != stride:
            if blk is not None:
                blk.close()
            self._block = blk = SharedHostBlock(B, stride, self.dst, self.group, register=register)
        return blk

    def _run_host_block(self, B, world, rank, b, copy):
        """Resident slice -> this rank's rows of the shared page-locked block (the kernel writes them
        over PCIe itself) -> barrier -> views on `dst`."""
        import torch
        import torch.distributed as dist

        ws = self._workspace(b)
        blk = self._shared_block(B, int(ws.row_stride))
        lo, _ = shard_range(B, world, rank)
        if b:
            ws.solve_rows_device(rows_ptr=blk.ptr + lo * blk.stride * 8)
        torch.cuda.current_stream().synchronize()       # this rank's rows are in host memory
        if world > 1:
            dist.barrier(self.group)
        if rank != self.dst:
            return None
        return self._host_solution(blk.rows[:B], copy)

    def _host_solution(self, h, copy):
        from .planner import HostSolution

        N = int(self.params.horizon)
        if copy:
            h = h.copy()
        if self.outputs == "all":
            return HostSolution.from_packed_rows(N, h)
        if self.outputs == "solution":
            return HostSolution.from_solution_rows(N, h, self.params)
        return HostSolution.from_control_rows(N, h)

    # -- the call ---------------------------------------------------------------------------
    def solve(self, p0, v0, goal, presliced: bool = False, global_B: Optional[int] = None,
              copy: bool = False):
        import time

        B, world, rank, p0, v0, goal = self._slice(p0, v0, goal, presliced, global_B)
        N = int(self.params.horizon)
        if not self.rows_supported():
            return self._solve_soa(p0, v0, goal, shard_counts(B, world), world, rank, N)
        if self.transport == "host_block":
            if self._rows_fn is not None:       # CPU plumbing tests: a stand-in writes the slice
                return self._run_host_block_standin(B, world, rank, p0, v0, goal, copy)
            if len(p0):
                self._workspace(len(p0)).set_inputs_device(p0, v0, goal)
            return self._run_host_block(B, world, rank, len(p0), copy)
        t0 = time.perf_counter()
        return self._finish(self._rows_local(p0, v0, goal), B, world, rank, copy, t0)

    def _run_host_block_standin(self, B, world, rank, p0, v0, goal, copy):
        import torch.distributed as dist

        rows = self._rows_fn(self.params, p0, v0, goal, self.outputs)
        rows = rows.numpy() if hasattr(rows, "numpy") else np.asarray(rows)
        blk = self._shared_block(B, int(rows.shape[1]), register=False)
        lo, _ = shard_range(B, world, rank)
        blk.rows[lo:lo + len(rows)] = rows
        if world > 1:
            dist.barrier(self.group)
        if rank != self.dst:
            return None
        return self._host_solution(blk.rows[:B], copy)

    def _finish(self, rows, B, world, rank, copy, t0):
        """gather -> pinned host block -> HostSolution views.  With t0 (solve()) the stages are
        timed with host clocks and synchronised in between; run() leaves everything queued on the
        current stream until the final copy has landed."""
        import time

        import torch

        from .planner import HostSolution

        counts = shard_counts(B, world)
        N = int(self.params.horizon)

        def sync(t):
            if t0 is not None and t is not None and t.is_cuda:
                torch.cuda.synchronize(t.device)
            return time.perf_counter()

        t1 = sync(rows)
        if world > 1:
            block, width = gather_row_blocks(rows, counts, self.dst, self.group, self._recv)
            self._recv = block
        else:
            block, width = rows, max(counts[0], 1)
        if rank != self.dst:
            return None
        t2 = sync(block)
        stride = block.shape[1]
        if block.is_cuda:
            if self._pinned is None or self._pinned.shape[0] < B or self._pinned.shape[1] != stride:
                self._pinned = torch.empty((max(B, 1), stride), dtype=torch.float64).pin_memory()
            host = self._pinned[:B]
            at = 0
            for r, c in enumerate(counts):      # one contiguous copy per rank slice
                if c:
                    host[at:at + c].copy_(block[r * width:r * width + c], non_blocking=True)
                at += c
            torch.cuda.current_stream(block.device).synchronize()
            h = host.numpy()
        else:
            h = np.concatenate([block[r * width:r * width + c].numpy() for r, c in enumerate(counts)], axis=0) \
                if world > 1 else block.numpy()
        t3 = time.perf_counter()
        if t0 is not None:
            self.last_timing = {"solve_ms": (t1 - t0) * 1e3, "gather_ms": (t2 - t1) * 1e3,
                                "to_host_ms": (t3 - t2) * 1e3, "row_bytes": int(stride * 8), "outputs": self.outputs}
        if copy:
            h = h.copy()
        if self.outputs == "all":
            return HostSolution.from_packed_rows(N, h)
        if self.outputs == "solution":
            return HostSolution.from_solution_rows(N, h, self.params)
        return HostSolution.from_control_rows(N, h)

    def _solve_soa(self, p0, v0, goal, counts, world, rank, N):
        """Stand-in / long-horizon path: SoA blocks, two padded gathers (rows do not fit N > 25)."""
        from .planner import HostSolution

        if self._solve_fn is not None:
            out, meta = self._solve_local(p0, v0, goal)
        else:
            ws = self._workspace(len(p0))
            b = len(p0)
            if b:
                ws.set_inputs_device(p0, v0, goal)
                sol = ws.solve_device()
                out, meta = sol.out[:, :b], sol.meta[:, :b]
            else:
                out, meta = ws.out[:, :0], ws.meta[:, :0]
        if world > 1:
            out = gather_columns(out, counts, self.dst, self.group)
            meta = gather_columns(meta, counts, self.dst, self.group)
        if rank != self.dst:
            return None
        if out.is_cuda:      # transpose on the device, one contiguous read-back
            return HostSolution.from_rows(N, out.t().contiguous().cpu().numpy(), meta.cpu().numpy())
        return HostSolution.from_blocks(N, out.cpu().numpy(), meta.cpu().numpy())
