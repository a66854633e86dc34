"""ctypes binding of the C ABI declared in include/glrgtv.h.

Only raw addresses cross this boundary (no torch types).  `load()` fails loudly when the nvcc-built
library is missing: there is no CPU fallback in the product path.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GLRGTV_LIB") or os.path.join(_PKG, "libglrgtv.so")   # GLRGTV_LIB: an alternative nvcc build

MAX_EDGES = 48
PAD_CLAMP, PAD_REFLECT = 0, 1
ABI_VERSION = 2

_STATUS = {
    0: "GLRGTV_OK", -1: "GLRGTV_ERR_SHAPE", -2: "GLRGTV_ERR_POINTER", -3: "GLRGTV_ERR_CUDA",
    -4: "GLRGTV_ERR_DEVICE", -5: "GLRGTV_ERR_WORKSPACE", -6: "GLRGTV_ERR_UNSUPPORTED",
}

fp = C.c_void_p  # every float* is passed as a raw address


class Shape(C.Structure):
    _fields_ = [("B", C.c_int32), ("G", C.c_int32), ("F", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


class Window(C.Structure):
    _fields_ = [("n_edges", C.c_int32), ("dh", C.c_int32 * MAX_EDGES), ("dw", C.c_int32 * MAX_EDGES)]


class Stats(C.Structure):
    _fields_ = [("p01", fp), ("p02a", fp), ("p02b", fp), ("p03", fp), ("n", C.c_int32), ("pad", C.c_int32)]


class OpParams(C.Structure):
    _fields_ = [("stats", Stats), ("multiM", fp)]


class BlockParams(C.Structure):
    _fields_ = [("gtv0", OpParams), ("glr0", OpParams), ("gtv1", OpParams), ("glr1", OpParams),
                ("alpha", fp), ("beta", fp), ("mu0", fp), ("ro0", fp), ("gamma0", fp),
                ("mu1", fp), ("ro1", fp), ("gamma1", fp), ("skip", fp)]


class BlockGrads(C.Structure):
    _fields_ = [(n, fp) for n in (
        "gtv0_stats", "glr0_stats", "gtv1_stats", "glr1_stats", "gtv0_M", "glr0_M", "gtv1_M", "glr1_M",
        "alpha", "beta", "mu0", "ro0", "gamma0", "mu1", "ro1", "gamma1", "skip")]


class BlockSaved(C.Structure):
    _fields_ = [(n, fp) for n in ("wT0", "wL0", "wT1", "wL1", "bA", "x1", "bB", "r1", "x2", "cT0", "cT1", "vc")]


def make_window(edges) -> Window:
    w = Window()
    if len(edges) > MAX_EDGES:
        raise ValueError(f"window with {len(edges)} edges exceeds GLRGTV_MAX_EDGES={MAX_EDGES}")
    w.n_edges = len(edges)
    for i, (dh, dw) in enumerate(edges):
        w.dh[i], w.dw[i] = int(dh), int(dw)
    return w


_P = C.POINTER
_SIGS = {
    # name: (restype, argtypes)
    "glrgtv_abi_version": (C.c_int, []),
    "glrgtv_last_cuda_error": (C.c_char_p, []),
    "glrgtv_check_device": (C.c_int, []),
    "glrgtv_launch_count": (C.c_ulonglong, []),
    "glrgtv_profile_enable": (C.c_int, [C.c_int]),
    "glrgtv_set_block_path": (C.c_int, [C.c_int]),
    "glrgtv_set_stream_loader": (C.c_int, [C.c_int]),
    "glrgtv_stream_launch_count": (C.c_ulonglong, []),
    "glrgtv_set_bwd_kernels": (C.c_int, [C.c_int]),
    "glrgtv_set_fwd_kernels": (C.c_int, [C.c_int]),
    "glrgtv_set_weights_kernels": (C.c_int, [C.c_int]),
    "glrgtv_set_proj_pipeline": (C.c_int, [C.c_int]),
    "glrgtv_weights_walk_launch_count": (C.c_ulonglong, []),
    "glrgtv_profile_read": (C.c_int, [_P(C.c_float), _P(C.c_int), C.c_int]),
    "glrgtv_edge_weights_fwd": (C.c_int, [_P(Shape), _P(Window), fp, fp, fp, fp]),
    "glrgtv_edge_weights_bwd": (C.c_int, [_P(Shape), _P(Window), fp, fp, fp, fp, fp, fp, fp, fp]),
    "glrgtv_normalize_fwd": (C.c_int, [_P(Shape), fp, fp, fp, fp]),
    "glrgtv_normalize_bwd": (C.c_int, [_P(Shape), fp, fp, fp, fp, fp, fp]),
    "glrgtv_gather_neighbors_fwd": (C.c_int, [_P(Shape), _P(Window), fp, fp, fp]),
    "glrgtv_gather_neighbors_bwd": (C.c_int, [_P(Shape), _P(Window), fp, fp, fp]),
    "glrgtv_stats_conv_fwd": (C.c_int, [_P(Shape), _P(Stats), fp, fp, fp]),
    "glrgtv_stats_conv_bwd": (C.c_int, [_P(Shape), _P(Stats), fp, fp, fp, fp, fp]),
    "glrgtv_stats_conv_t_fwd": (C.c_int, [_P(Shape), _P(Stats), fp, fp, fp]),
    "glrgtv_stats_conv_t_bwd": (C.c_int, [_P(Shape), _P(Stats), fp, fp, fp, fp, fp]),
    "glrgtv_op_L_fwd": (C.c_int, [_P(Shape), _P(Window), fp, fp, fp, fp]),
    "glrgtv_op_L_bwd": (C.c_int, [_P(Shape), _P(Window), fp, fp, fp, fp, fp, fp]),
    "glrgtv_op_C_fwd": (C.c_int, [_P(Shape), _P(Window), fp, fp, fp, fp]),
    "glrgtv_op_C_bwd": (C.c_int, [_P(Shape), _P(Window), fp, fp, fp, fp, fp, fp]),
    "glrgtv_op_Ct_fwd": (C.c_int, [_P(Shape), _P(Window), fp, fp, fp, fp]),
    "glrgtv_op_Ct_bwd": (C.c_int, [_P(Shape), _P(Window), fp, fp, fp, fp, fp, fp]),
    "glrgtv_soft_threshold_fwd": (C.c_int, [_P(Shape), C.c_int, fp, fp, fp, fp]),
    "glrgtv_soft_threshold_bwd": (C.c_int, [_P(Shape), C.c_int, fp, fp, fp, fp, fp, fp]),
    "glrgtv_mixture_fwd": (C.c_int, [_P(Shape), fp, fp, fp, fp]),
    "glrgtv_mixture_bwd": (C.c_int, [_P(Shape), fp, fp, fp, fp, fp, fp]),
    "glrgtv_pool2_fwd": (C.c_int, [_P(Shape), fp, fp, fp]),
    "glrgtv_unpool2_fwd": (C.c_int, [_P(Shape), fp, fp, fp]),
    "glrgtv_proj_gemm_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "glrgtv_proj_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, C.c_size_t, fp]),
    "glrgtv_space_to_depth": (C.c_int, [C.c_int, C.c_long, C.c_int, C.c_int, fp, fp, fp]),
    "glrgtv_pixel_rstd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_long, C.c_float, fp, fp, fp]),
    "glrgtv_dwconv_gate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, fp, fp, fp]),
    "glrgtv_dwconv_gate_bwd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, fp, fp, fp, fp]),
    "glrgtv_pixel_norm_bwd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_long, fp, fp, fp, fp, fp, fp, fp]),
    "glrgtv_proj_wgrad": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp]),
    "glrgtv_block_fwd": (C.c_int, [_P(Shape), _P(BlockParams), fp, fp, fp, fp, _P(BlockSaved), fp]),
    "glrgtv_block_fwd_stage": (C.c_int, [C.c_int, _P(Shape), _P(BlockParams), fp, fp, fp, fp, _P(BlockSaved), C.c_int, C.c_int, fp]),
    "glrgtv_block_bwd_workspace_bytes": (C.c_size_t, [_P(Shape)]),
    "glrgtv_block_bwd": (C.c_int, [_P(Shape), _P(BlockParams), fp, fp, fp, _P(BlockSaved), fp, fp, fp, fp,
                                   _P(BlockGrads), fp, C.c_size_t, fp]),
}
EXPORTED = tuple(_SIGS)


def bind(lib, only=None):
    """attach argtypes / restypes; raises AttributeError if a declared symbol is not exported."""
    for name, (res, args) in _SIGS.items():
        if only is not None and name not in only:
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def load(path: str = None):
    """Load and bind libglrgtv.so (built by build.py).  No fallback: raises if it is not there."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: build it with `python -m imagerestoration_development_unrolling_b200.build` "
            "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
    lib = bind(C.CDLL(p))
    v = lib.glrgtv_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"libglrgtv ABI {v} != binding ABI {ABI_VERSION}: rebuild the library")
    if path is None:
        _lib = lib
    return lib


def check(rc: int, lib=None, what: str = ""):
    if rc == 0:
        return
    msg = _STATUS.get(rc, str(rc))
    if rc == -3 and lib is not None:
        msg += ": " + (lib.glrgtv_last_cuda_error() or b"").decode()
    raise RuntimeError(f"libglrgtv {what} failed: {msg}")


def call(lib, name, *args):
    """Marshal one C-ABI call: Structures go by reference, anything with .data_ptr() as its address."""
    cargs = []
    for a in args:
        if isinstance(a, C.Structure):
            cargs.append(C.byref(a))
        elif a is None:
            cargs.append(None)
        elif hasattr(a, "data_ptr"):
            cargs.append(a.data_ptr())
        else:
            cargs.append(a)
    rc = getattr(lib, name)(*cargs)
    check(rc, lib, name)


def make_shape(B, G, F, H, W) -> Shape:
    return Shape(int(B), int(G), int(F), int(H), int(W))


def make_stats(p1, pa, pb, p3, pad=PAD_CLAMP) -> Stats:
    n = p1.numel()
    return Stats(p1.data_ptr(), pa.data_ptr(), pb.data_ptr(), p3.data_ptr(), int(n), int(pad))
