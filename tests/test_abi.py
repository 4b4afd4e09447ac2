"""CPU: the C-ABI library loads, exports exactly what include/cafe_gpu.h declares, and refuses to run
without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from cafe_b200 import buildlib
from cafe_b200 import gpu as cgpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "cafe_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cafe_gpu_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) >= 25
    assert sorted(cgpu.ABI_SYMBOLS) == syms


def test_library_exports_every_declared_symbol():
    L = cgpu.load_library()
    for s in header_symbols():
        assert hasattr(L, s), s
    assert L.cafe_gpu_abi_version() == 2


def test_signatures_are_plain_c():
    # no torch / C++ types at the boundary: the header must compile as C
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write('#include "cafe_gpu.h"\nint main(void){return cafe_gpu_abi_version()==CAFE_GPU_ABI_VERSION?0:1;}\n')
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", src, "-o", os.path.join(td, "t.o")],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = cgpu.load_library()
    h = C.c_void_p()
    rc = L.cafe_gpu_create(C.byref(h), -1)
    assert rc == -1 and not h.value
    assert b"no CPU fallback" in L.cafe_gpu_last_error(None)
    with pytest.raises(cgpu.CafeGpuError, match="no CPU fallback"):
        cgpu.CafeGpu()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cafe_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".h", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "cafe_oracle" not in txt and "libcafe_ref" not in txt, f


def test_built_artifacts_exist():
    assert os.path.exists(buildlib.GPU_LIB) and os.path.exists(buildlib.HOST_LIB) and os.path.exists(buildlib.SHELL_BIN)
