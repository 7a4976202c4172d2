"""Builds libb200backend.so (the C-ABI library) with nvcc for sm_100a, in-tree.

``python candle-vllm_b200/build.py [--force] [--verbose]``.  Each ``csrc/*.cu`` is compiled to an
object (parallel), then linked into ``candle-vllm_b200/libb200backend.so`` (static cudart, so the
library has no run-time dependency beyond the driver).
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libb200backend.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CCBIN = "/usr/bin/g++"

FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-ccbin", CCBIN, "-Xcompiler", "-fPIC,-fvisibility=default", "--expt-relaxed-constexpr",
         "-I", os.path.join(HERE, "..", "include")]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        if force or _newer(o, [s] + hdrs):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append((s, cmd))

    def run(job):
        s, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for s, r in ex.map(run, jobs):
                if verbose or r.returncode:
                    sys.stderr.write(f"== {os.path.basename(s)}\n{r.stdout}{r.stderr}\n")
                if r.returncode:
                    raise RuntimeError(f"nvcc failed on {s}")
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if force or jobs or _newer(SO, objs):
        cmd = [NVCC, "-shared", "-o", SO, "-ccbin", CCBIN, "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static"] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
