"""Build libdinoseg.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The shared object lives at dino_b200/lib/libdinoseg.so: git-ignored, but it travels to the
GPU box with the repo snapshot.  `python -m dino_b200.build` rebuilds it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdinoseg.so")
SOURCES = [os.path.join(CSRC, "dinoseg_api.cu")]
def headers() -> list:
    """Everything the translation unit includes: every csrc/*.cuh plus the public header (globbed, so that a new
    kernel header can never be forgotten by the staleness check)."""
    import glob
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(os.path.dirname(HERE), "include", "dinoseg.h")]


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libdinoseg.so")
    return nvcc


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + headers())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if it is missing or older than its sources. Returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
