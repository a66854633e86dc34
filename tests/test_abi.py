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


def test_header_is_plain_c_and_struct_layouts_match_the_binding(tmp_path):
    """include/glrgtv.h compiles as strict C99 (the boundary is a C ABI, not C++), and every POD struct has the size and field
    offsets the ctypes binding assumes - a drift between header and binding would corrupt arguments silently"""
    pairs = [("glrgtv_shape", L.Shape), ("glrgtv_window", L.Window), ("glrgtv_stats", L.Stats), ("glrgtv_opparams", L.OpParams),
             ("glrgtv_block_params", L.BlockParams), ("glrgtv_block_grads", L.BlockGrads), ("glrgtv_block_saved", L.BlockSaved)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "glrgtv.h"', "int main(void) {"]
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{field} %zu\\n", offsetof({cname}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in pairs:
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for field, _ in cls._fields_:
            assert int(got[f"{cname}.{field}"]) == getattr(cls, field).offset, (cname, field)
