#!/usr/bin/env python
"""Builds an experimental variant of the library next to the product build, for A/B runs on the
GPU box (tools/ab_libs.py, DART_SE3MPC_LIB=...):

  python tools/build_variant.py TAG [-DFOO ...] [--units a.cu,b.cu]

compiles every unit with the extra flags into /tmp/dart_variant_TAG/ and links
gpurun_scratch/libdart_TAG.so (git-ignored; travels to the box with the snapshot)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dart_planner_b200 import build as b  # noqa: E402


def main():
    tag = sys.argv[1]
    extra = [a for a in sys.argv[2:] if not a.startswith("--units")]
    out = os.path.join("/tmp", "dart_variant_" + tag)      # objects stay out of the gpurun snapshot
    os.makedirs(out, exist_ok=True)
    os.makedirs(os.path.join(ROOT, "gpurun_scratch"), exist_ok=True)
    lib = os.path.join(ROOT, "gpurun_scratch", f"libdart_{tag}.so")

    def one(unit):
        obj = os.path.join(out, unit.replace(".cu", ".o"))
        cmd = [b.nvcc()] + b.ARCH + b.COMMON + b.UNITS[unit] + extra + ["-Xptxas", "-v", "-c", os.path.join(b.CSRC, unit), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {unit}:\n{r.stderr[-4000:]}")
        return obj, r.stderr

    with ThreadPoolExecutor(len(b.UNITS)) as ex:
        res = list(ex.map(one, b.UNITS))
    with open(os.path.join(out, "ptxas.log"), "w") as fh:
        fh.write("\n".join(l for _, l in res))
    subprocess.run([b.nvcc()] + b.ARCH + ["-shared", "-o", lib] + [o for o, _ in res], check=True)
    print(lib)


if __name__ == "__main__":
    main()
