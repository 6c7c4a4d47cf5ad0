"""In-tree build of libstepsb200.so (nvcc, sm_100a only) and of the C++ drop-in shim objects."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _stale(target: str, sources: list) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library(force: bool = False, verbose: bool = False) -> str:
    csrc = os.path.join(HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "steps_b200.h"))
    out = os.path.join(HERE, "libstepsb200.so")
    if force or _stale(out, srcs):
        cmd = [NVCC, "-ccbin", HOST_CXX, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
               "-cudart", "static", "-o", out, os.path.join(csrc, "engine.cu"), "-ldl"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        print("[steps_b200.build]", " ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
