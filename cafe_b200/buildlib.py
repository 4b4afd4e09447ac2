"""Build helpers: compile the sm_100a C-ABI library and the C++ host library in-tree."""
from __future__ import annotations

import fcntl
import glob
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
GPU_LIB = os.environ.get("CAFE_GPU_LIB") or os.path.join(HERE, "libcafe_gpu.so")  # CAFE_GPU_LIB: an A/B build (tools/k2_variants.sh)
HOST_LIB = os.path.join(HERE, "libcafe_host.so")
SHELL_BIN = os.path.join(HERE, "cafe_gpu_shell")
STAMP = os.path.join(HERE, "build", ".sources.sha1")


def _source_hash() -> str:
    h = hashlib.sha1()
    pats = ["csrc/*.cu", "csrc/*.cuh", "csrc/Makefile", "host/*.cpp", "host/*.h", "host/Makefile", "../include/*.h"]
    for pat in pats:
        for f in sorted(glob.glob(os.path.join(HERE, pat))):
            h.update(os.path.basename(f).encode())
            with open(f, "rb") as fp:
                h.update(fp.read())
    return h.hexdigest()


def build(verbose: bool = False) -> None:
    """nvcc -gencode arch=compute_100a,code=sm_100a (csrc/Makefile) + g++ (host/Makefile)."""
    out = None if verbose else subprocess.DEVNULL
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)  # one builder at a time (torchrun starts one process per GPU)
        # -z defs: an undefined symbol fails the link of the library here, not the load on the GPU box
        subprocess.run(["make", "-j8", "-C", os.path.join(HERE, "csrc")], check=True, stdout=out)
        subprocess.run(["make", "-j8", "-C", os.path.join(HERE, "host")], check=True, stdout=out)
        with open(STAMP, "w") as fp:
            fp.write(_source_hash())


def ensure_built() -> None:
    """Rebuild when any source differs from what the libraries were built from (a content hash, so that copying the tree
    to another machine — which changes every mtime — does not trigger a build there)."""
    try:
        with open(STAMP) as fp:
            fresh = fp.read().strip() == _source_hash()
    except OSError:
        fresh = False
    if fresh and os.path.exists(GPU_LIB) and os.path.exists(HOST_LIB):
        return
    build()
