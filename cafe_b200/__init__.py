"""cafe_b200 — B200-native implementation of CAFE's per-family birth–death likelihood hot path.

Layout
  csrc/   hand-written sm_100a CUDA kernels + the C-ABI of include/cafe_gpu.h  -> libcafe_gpu.so
  host/   C++ mirror of the reference's entry points above the ABI              -> libcafe_host.so
  gpu.py / host.py   ctypes bindings (plumbing for tests and bench.py)
  synth.py           synthetic in-model family tables for the benchmark configurations
  sharding.py        family sharding across ranks (one process per GPU, torch.distributed)

There is no CPU fallback: the product path fails loudly when the CUDA library or a device is missing.
"""
from .buildlib import build  # noqa: F401

__all__ = ["build"]
