"""Build helpers: compile the sm_100a C-ABI library and the C++ host library in-tree."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
GPU_LIB = os.path.join(HERE, "libcafe_gpu.so")
HOST_LIB = os.path.join(HERE, "libcafe_host.so")
SHELL_BIN = os.path.join(HERE, "cafe_gpu_shell")


def build(verbose: bool = False) -> None:
    """nvcc -gencode arch=compute_100a,code=sm_100a (csrc/Makefile) + g++ (host/Makefile)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.run(["make", "-j8", "-C", os.path.join(HERE, "csrc")], check=True, stdout=out)
    subprocess.run(["make", "-j8", "-C", os.path.join(HERE, "host")], check=True, stdout=out)


def ensure_built() -> None:
    if not (os.path.exists(GPU_LIB) and os.path.exists(HOST_LIB)):
        build()
