"""The C-ABI library: builds, loads, and exports every symbol include/glrgtv.h declares (no compute, CPU only)."""
import ctypes
import os
import re
import subprocess

from imagerestoration_development_unrolling_b200 import _lib as L
from imagerestoration_development_unrolling_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "glrgtv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(glrgtv_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(L.EXPORTED)


def test_cuda_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(B.build_cuda())     # nvcc cross-compiles for sm_100a without a GPU
    for name in _declared():
        assert hasattr(lib, name), name
    L.bind(lib)
    assert lib.glrgtv_abi_version() == L.ABI_VERSION


def test_cuda_library_is_sm100a_only():
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", B.build_cuda()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_validation_without_gpu():
    """shape / pointer errors are reported before anything is launched"""
    lib = L.bind(ctypes.CDLL(B.build_cuda()))
    shp = L.make_shape(1, 1, 1, 3, 4)           # odd H: the block needs even sizes
    assert lib.glrgtv_pool2_fwd(ctypes.byref(shp), 16, 16, None) == -1
    shp = L.make_shape(1, 1, 1, 4, 4)
    assert lib.glrgtv_pool2_fwd(ctypes.byref(shp), None, 16, None) == -2
    assert lib.glrgtv_pool2_fwd(ctypes.byref(shp), 18, 16, None) == -2      # misaligned


def test_missing_library_fails_loudly(tmp_path):
    import pytest
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        L.load(str(tmp_path / "nope.so"))
