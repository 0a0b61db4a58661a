"""Builds libleela_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libleela_b200.so")
SOURCES = ["lb2_api.cu", "lb2_kernels.cu", "lb2_planes.cpp"]
HEADERS = ["lb2_kernels.cuh", "lb2_ptx.cuh", os.path.join("..", "..", "include", "leela_b200.h")]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB, csrc: str = CSRC) -> str:
    """`defines`, `out`, `csrc` exist for tools/ab_variants.py (tuning builds side by side); the product is LIB."""
    if out == LIB and not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", out] + [f"-D{d}" for d in defines] + \
          [os.path.join(csrc, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
