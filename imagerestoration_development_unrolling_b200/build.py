"""Build libglrgtv.so in-tree with nvcc for sm_100a (and, for the tests, the g++ emulation build).

    python -m imagerestoration_development_unrolling_b200.build          # the product library
    python -m imagerestoration_development_unrolling_b200.build --emu    # tests/emu/libglrgtv_emu.so

The .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libglrgtv.so")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libglrgtv_emu.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps():
    return _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))


def build_cuda(force=False, verbose=False):
    """nvcc -> libglrgtv.so.  Object files are kept under csrc/build so edits recompile one file."""
    if not force and not _stale(LIB, _deps()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [d for d in _deps() if not d.endswith(".cu")]
    objs, procs = [], []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


def build_emu(force=False, asan=False):
    """g++ emulation build of the same kernel sources (tests only; never loaded by the package).
    asan=True: a second library built with AddressSanitizer and exact-size shared-memory blocks (tools/emu_asan.sh)."""
    target = EMU_LIB[:-3] + "_asan.so" if asan else EMU_LIB
    if not force and not _stale(target, _deps()):
        return target
    os.makedirs(os.path.dirname(target), exist_ok=True)
    opt = ["-O1", "-g", "-fno-omit-frame-pointer", "-fsanitize=address", "-DGLRGTV_EMU_EXACT_SMEM"] if asan else ["-O2"]
    cmd = ["g++"] + opt + ["-std=c++17", "-fPIC", "-shared", "-DGLRGTV_EMU", "-Wall", "-Wno-unused-function",
                           "-Wno-unknown-pragmas", "-Wno-unused-variable", "-o", target]
    for src in _sources():
        cmd += ["-x", "c++", src]
    subprocess.check_call(cmd)
    return target


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu(force=True, asan="--asan" in sys.argv))
    else:
        print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
