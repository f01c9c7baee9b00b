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
    """Compile every ``csrc/*.cu`` to an object (only the stale ones unless ``force``; objects of A/B builds with extra flags
    live in their own directory) and link the shared library.  The link goes to a temporary file that is renamed into place,
    under a lock, so that concurrent ranks of a fresh checkout never ``dlopen`` a half-written library."""
    if not force and not needs_build():
        return LIB_PATH
    import fcntl
    import hashlib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = os.environ.get("HERMNET_B200_NVCC_FLAGS", "").split()      # e.g. -DHN_FWD_PIPE=0 for A/B builds
    tag = hashlib.sha1(" ".join(extra).encode()).hexdigest()[:8] if extra else "default"
    obj_dir = os.path.join(HERE, "build", tag)
    os.makedirs(obj_dir, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():      # another rank built it while we waited
            return LIB_PATH
        headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h"))
        hdr_t = max(os.path.getmtime(h) for h in headers)
        objs, procs = [], []
        for src in sources():
            obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
            objs.append(obj)
            if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
                cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas", "-v"] if verbose else []) + \
                    ["-c", "-o", obj, src]
                procs.append((cmd, subprocess.Popen(cmd)))
        for cmd, pr in procs:
            if pr.wait() != 0:
                raise subprocess.CalledProcessError(pr.returncode, cmd)
        tmp = LIB_PATH + f".tmp{os.getpid()}"
        subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs)
        os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
