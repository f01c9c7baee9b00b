"""In-tree build of the CUDA library: ``nvcc -gencode arch=compute_100a,code=sm_100a`` of ``csrc/*.cu`` into
``hermnet_b200/lib/libhermnet_b200.so`` (git-ignored; travels to the GPU box with the snapshot).

    python -m hermnet_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
# HERMNET_B200_LIB: load / build another file instead (A/B builds made with HERMNET_B200_NVCC_FLAGS)
LIB_PATH = os.environ.get("HERMNET_B200_LIB") or os.path.join(LIB_DIR, "libhermnet_b200.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "549"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = os.environ.get("HERMNET_B200_NVCC_FLAGS", "").split()      # e.g. -DHN_FWD_PIPE=0 for A/B builds
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
