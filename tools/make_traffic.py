#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` captures of the solve kernel (one launch each):
    python tools/make_traffic.py ROUND B=rep.ncu-rep [B=rep.ncu-rep ...]   (N = 8, bench distribution)
Per batch size: DRAM bytes of the launch (read + write), executed warp instructions, the issue
rate and the FP64 pipe's active share as ncu measured them.  bench.py copies these into
`roofline.traffic` / `roofline.issue`; they are measured under the profiler (cold caches,
serialised launch), so only per-launch COUNTS are used, never the durations."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "warp_inst", "sm__inst_executed.sum": "warp_inst_sm",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers", "gpu__time_duration.sum": "duration_under_ncu",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            x = float(v.replace(",", ""))
            d[WANT[h]] = x * UNIT_SCALE.get(u, 1.0) if "bytes" in h else x
            if h == "gpu__time_duration.sum":
                d["duration_unit"] = u
    return d


def main():
    rnd = sys.argv[1]
    N = 8
    res = {"source": f"ncu --set full --clock-control none, one launch of se3mpc_solve_kernel per batch size "
                     f"(profiles/{rnd}_solve_kernel_B*_details.txt); dram__bytes_read.sum + dram__bytes_write.sum, "
                     f"smsp__inst_executed.sum, smsp__issue_active, sm__pipe_fp64_cycles_active"}
    for a in sys.argv[2:]:
        B, rep = a.split("=", 1)
        B = int(B)
        d = read(rep)
        e = {"dram_bytes_per_launch": d["dram_read"] + d["dram_write"], "dram_read": d["dram_read"],
             "dram_write": d["dram_write"], "alg_bytes": (8 * 9 + 8 * (19 * N + 1) + 16) * B,
             "warp_inst_per_launch": d.get("warp_inst"), "warp_inst_per_solve": d.get("warp_inst", 0) / B}
        for k in ("issue_active_pct", "fp64_pipe_active_pct", "threads_per_inst", "achieved_occupancy_pct",
                  "registers", "l1_hit_pct", "l2_hit_pct"):
            if k in d:
                e[k] = d[k]
        res[f"B{B}_N{N}"] = e
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
