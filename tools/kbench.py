#!/usr/bin/env python
"""Kernel tuning bench: times the solve kernel for each DART_SE3MPC_VARIANT index on the bench
distribution and checks every variant's result against variant-default (same arithmetic for the
same lane count).  usage: python tools/kbench.py [variants=default,5,6,...] [B list] [N]
DART_KBENCH_NOFLUSH=1 skips the L2 flush between launches (warm instruction / data caches)."""
import ctypes as C
import os
import statistics
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dart_planner_b200 as dp  # noqa: E402
from dart_planner_b200 import _cabi  # noqa: E402
from dart_planner_b200.config import make_params  # noqa: E402
from dart_planner_b200.planner import BatchWorkspace  # noqa: E402

variants = sys.argv[1].split(",") if len(sys.argv) > 1 else ["default"]
Bs = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4096, 65536]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 8
L = _cabi.lib()
params = make_params(dp.SE3MPCConfig(prediction_horizon=N, dt=0.1),
                     max_corrections=int(os.environ.get("DART_KBENCH_M", "10")))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
for B in Bs:
    rng = np.random.default_rng(1)
    p0 = rng.uniform(-10, 10, (B, 3))
    goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    ws = BatchWorkspace(params, B, pinned=False)
    ws.set_inputs_device(p0, np.zeros((B, 3)), goal)
    base = None
    for v in variants:
        if v == "default":
            os.environ.pop("DART_SE3MPC_VARIANT", None)
        else:
            os.environ["DART_SE3MPC_VARIANT"] = v
        info = [C.c_int32() for _ in range(5)]
        L.dart_se3mpc_kernel_info(C.byref(params), B, *[C.byref(i) for i in info])
        reps = 20 if B <= 65536 else 5
        for _ in range(3):
            ws.solve_device(stream)
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            if os.environ.get("DART_KBENCH_NOFLUSH"):
                torch.cuda._sleep(400000)   # keeps the GPU busy while the launch is queued; L2 stays warm
            else:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            ws.solve_device(stream)
            b.record(stream)
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        out = ws.out[:, :B].clone()
        meta = ws.meta[:, :B].clone()
        if base is None:
            base = (out, meta)
            agree = "base"
        else:
            agree = f"max|dx|={float((out - base[0]).abs().max()):.2e} meta_eq={bool((meta == base[1]).all())}"
        med = statistics.median(ms)
        print(f"B={B:8d} N={N} variant={v:8s} lanes={info[0].value} block={info[1].value} grid={info[2].value} "
              f"smem={info[3].value} regs={info[4].value}  {med * 1e3:9.1f} us  {B / med / 1e3:8.2f} Msolves/s  {agree}",
              flush=True)
