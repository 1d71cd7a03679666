"""Builds libnode4j_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnode4j_b200.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "solver_kernels.cuh", "layer_kernels.cuh", "tc_common.cuh", "conv_tc.cuh", "conv_wgrad_tc.cuh", "gemm_tc.cuh", "mlp_small.cuh", "outer_net.cuh", "outer_net_tc.cuh", "engine.cuh", "../../include/node4j.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--fmad=true",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Concurrent callers (torchrun ranks importing the package at once) serialise on a lock file; the library is written to a
    temporary name and renamed into place, so nobody ever dlopens a half-written file."""
    if not force and not _stale():
        return LIB
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():      # another process built it while this one waited
                return LIB
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            tmp = f"{LIB}.{os.getpid()}.tmp"
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ))
            if verbose:
                sys.stderr.write(r.stderr)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
            os.replace(tmp, LIB)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
