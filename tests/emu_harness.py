"""Test-only harness: drives the g++ EMULATION build of the kernel sources (tests/emu/libglrgtv_emu.so)
with CPU tensors, so index arithmetic and border rules can be checked against the oracle without a GPU.
Never imported by the product package."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from imagerestoration_development_unrolling_b200 import _lib as L  # noqa: E402
from imagerestoration_development_unrolling_b200 import build as B  # noqa: E402

_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        path = B.build_emu(asan=os.environ.get("GLRGTV_EMU_ASAN") == "1")     # tools/emu_asan.sh preloads libasan and sets this
        _emu = L.bind(ctypes.CDLL(path), only=[n for n in L.EXPORTED if hasattr(ctypes.CDLL(path), n)])
    return _emu


def call(name, *args):
    L.call(emu_lib(), name, *args)
