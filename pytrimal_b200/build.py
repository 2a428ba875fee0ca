"""Build libtrimal_cuda.so in-tree with nvcc for sm_100a (no other arch)."""
from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libtrimal_cuda.so")
SOURCES = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
HEADERS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh")) +
                 glob.glob(os.path.join(HERE, "csrc", "*.h"))) + [
    os.path.join(ROOT, "include", "trimal_cuda.h")
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-ldl",
]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile pytrimal_b200/csrc/*.cu -> pytrimal_b200/libtrimal_cuda.so."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "csrc")]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB_PATH, *SOURCES]
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
