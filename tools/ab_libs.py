#!/usr/bin/env python
"""A/B of two builds of the library in ONE process on one box (box-to-box variation is larger than
most kernel changes): interleaved event-timed launches with an L2 flush, plus 20 launches back to
back between one event pair (the event tick is ~1 us), and a bit-for-bit comparison of the results.
usage: python tools/ab_libs.py libA.so libB.so [libC.so ...] [B list]     (N = 8, bench distribution, seed 1)"""
import ctypes as C, statistics, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, torch
import dart_planner_b200 as dp
from dart_planner_b200 import _cabi
from dart_planner_b200.config import make_params
libs = {}
paths = [a for a in sys.argv[1:] if a.endswith('.so')]
rest = [a for a in sys.argv[1:] if not a.endswith('.so')]
for path in paths:
    L = C.CDLL(path)
    vp, i64 = C.c_void_p, C.c_int64
    L.dart_se3mpc_solve_batch.argtypes = [C.POINTER(_cabi.Params), i64, i64] + [vp] * 16 + [vp]
    libs[path] = L
Bs = [int(b) for b in rest[0].split(",")] if rest else [4096, 65536, 1 << 20]
N = 8
params = make_params(dp.SE3MPCConfig(prediction_horizon=N, dt=0.1))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
for B in Bs:
    rng = np.random.default_rng(1)
    p0 = rng.uniform(-10, 10, (B, 3)); goal = np.concatenate([rng.uniform(-15, 15, (B, 2)), rng.uniform(3, 8, (B, 1))], axis=1)
    inp = torch.zeros((9, B), dtype=torch.float64, device="cuda")
    inp[0:3] = torch.as_tensor(p0.T).cuda(); inp[6:9] = torch.as_tensor(goal.T).cuda()
    outs = {}
    res = {k: [] for k in libs}
    def run(L, out, meta):
        es = 8 * B; b = out.data_ptr(); m = meta.data_ptr(); i = inp.data_ptr()
        rc = L.dart_se3mpc_solve_batch(C.byref(params), B, B, i, i + 3 * es, i + 6 * es, None, None, None, b, b + 72 * es,
                                      m, m + 4 * B, m + 8 * B, m + 12 * B, b + 73 * es, b + 97 * es, b + 121 * es, b + 145 * es, stream.cuda_stream)
        assert rc == 0
    for k, L in libs.items():
        outs[k] = (torch.zeros((153, B), dtype=torch.float64, device="cuda"), torch.zeros((4, B), dtype=torch.int32, device="cuda"))
        for _ in range(3): run(L, *outs[k])
    torch.cuda.synchronize()
    for rep in range(6):
        for k, L in libs.items():          # interleaved
            for _ in range(5 if B <= 65536 else 2):
                flush.zero_(); a = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
                a.record(stream); run(L, *outs[k]); e.record(stream); torch.cuda.synchronize(); res[k].append(a.elapsed_time(e))
    # finer: 20 launches back to back between two events (no flush; the event tick is ~1 us)
    chain = {}
    for rep in range(5):
        for k, L in libs.items():
            torch.cuda._sleep(2000000)
            a = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(20): run(L, *outs[k])
            e.record(stream); torch.cuda.synchronize(); chain.setdefault(k, []).append(a.elapsed_time(e) / 20)
    print("   chained x20: " + " | ".join(f"{k.split('/')[-1]} {statistics.median(v)*1e3:.2f} us" for k, v in chain.items()))
    ks = list(libs)
    same = all(bool((outs[ks[0]][0] == outs[k][0]).all()) and bool((outs[ks[0]][1] == outs[k][1]).all()) for k in ks[1:])
    for k in ks[1:]:
        dx = float((outs[ks[0]][0][:72] - outs[k][0][:72]).abs().max())
        ndiff = int((outs[ks[0]][0][:72] != outs[k][0][:72]).any(dim=0).sum())
        print(f"   {k.split('/')[-1]} vs {ks[0].split('/')[-1]}: {ndiff} problems differ in x, max |dx| {dx:.2e}, "
              f"counters equal: {bool((outs[ks[0]][1] == outs[k][1]).all())}")
    print(f"B={B}: " + " | ".join(f"{k.split('/')[-1]} {statistics.median(v)*1e3:.1f} us ({B/statistics.median(v)/1e3:.1f} M/s)" for k, v in res.items()) + f" identical={same}", flush=True)
