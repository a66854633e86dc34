"""torch.library custom ops over the C ABI (include/glrgtv.h).

Every op launches hand-written sm_100a kernels from libglrgtv.so on the caller's current CUDA stream.
There is NO fallback: CPU tensors, a missing library or a non-Blackwell device raise.  The ops are
registered with fake (meta) implementations and autograd formulas so that `nn.Module.compile()`
(which every reference script calls, scripts_v2/run_abtract_lightformer_GGTV_GGLR_sigma25.py:130)
treats them as opaque nodes.

Operator <-> reference method (V1X0 = deep_multiscale_GGLR_GGTV_v1x0.py):
  edge_weights      GLRFast/GTVFast.extract_edge_weights   V1X0:160-175, 391-407
  stats_conv(_t)    stats_conv / stats_conv_transpose      V1X0:177-215
  op_L              GLRFast.op_L_norm                      V1X0:218-228
  op_C / op_Ct      GTVFast.op_C / op_C_transpose (cores)  V1X0:452-516
  soft_threshold    MixtureGTVGLR.soft_threshold           V1X0:684-704
  pool2 / unpool2   the 0.25 depthwise 2x2 (transposed) conv  V1X0:613, 662-679
  lowpass_block     LocalLowpassFilteringBlock.forward     V1X0:707-811, 985-988
"""
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib as L

_NS = "glrgtv"
launch_count = 0  # C-ABI calls issued (each is >= 1 kernel launch); bench.py reports kernel launches itself


def _lib():
    return L.load()


def _stream(t: Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _chk(*ts: Tensor):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("glrgtv ops run on CUDA tensors only (sm_100a kernels; no CPU fallback)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"glrgtv ops are float32, got {t.dtype}")


def _c(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


_checked_devices = set()


def _call(name: str, ref: Tensor, *args):
    global launch_count
    launch_count += 1
    with torch.cuda.device(ref.device):
        if ref.device.index not in _checked_devices:      # first use of this device: the kernels are sm_100a only
            L.check(_lib().glrgtv_check_device(), _lib(), f"check_device({ref.device})")
            _checked_devices.add(ref.device.index)
        L.call(_lib(), name, *args, _stream(ref))


def _win(edges: Sequence[int]) -> L.Window:
    return L.make_window([(edges[2 * i], edges[2 * i + 1]) for i in range(len(edges) // 2)])


def _shape5(x: Tensor) -> L.Shape:
    B, G, F, H, W = x.shape
    return L.make_shape(B, G, F, H, W)


def flat_edges(edges) -> List[int]:
    out = []
    for dh, dw in edges:
        out += [int(dh), int(dw)]
    return out


# ====================================================================== edge weights
@torch.library.custom_op(f"{_NS}::edge_weights", mutates_args=())
def edge_weights(feat: Tensor, multiM: Tensor, edges: List[int]) -> Tensor:
    _chk(feat, multiM)
    feat, multiM = _c(feat), _c(multiM)
    B, G, F, H, W = feat.shape
    w = feat.new_empty(B, G, len(edges) // 2, H, W)
    _call("glrgtv_edge_weights_fwd", feat, _shape5(feat), _win(edges), feat, multiM, w)
    return w


@edge_weights.register_fake
def _(feat, multiM, edges):
    B, G, F, H, W = feat.shape
    return feat.new_empty(B, G, len(edges) // 2, H, W)


@torch.library.custom_op(f"{_NS}::edge_weights_bwd", mutates_args=())
def edge_weights_bwd(feat: Tensor, multiM: Tensor, w: Tensor, gw: Tensor, edges: List[int]) -> Tuple[Tensor, Tensor]:
    _chk(feat, multiM, w, gw)
    feat, multiM, w, gw = _c(feat), _c(multiM), _c(w), _c(gw)
    B, G, F, H, W = feat.shape
    E = len(edges) // 2
    gfeat, gM = torch.empty_like(feat), torch.zeros_like(multiM)
    scratch = feat.new_empty(B * G * (E + 1) * H * W)
    _call("glrgtv_edge_weights_bwd", feat, _shape5(feat), _win(edges), feat, multiM, w, gw, gfeat, gM, scratch)
    return gfeat, gM


@edge_weights_bwd.register_fake
def _(feat, multiM, w, gw, edges):
    return torch.empty_like(feat), torch.empty_like(multiM)


def _ew_setup(ctx, inputs, output):
    feat, multiM, edges = inputs
    ctx.save_for_backward(feat, multiM, output)
    ctx.edges = edges


def _ew_bwd(ctx, gw):
    feat, multiM, w = ctx.saved_tensors
    gfeat, gM = edge_weights_bwd(feat, multiM, w, gw, ctx.edges)
    return gfeat, gM, None


edge_weights.register_autograd(_ew_bwd, setup_context=_ew_setup)


# ====================================================================== stats conv / transpose
def _stats_struct(ps: Sequence[Tensor], pad: int) -> L.Stats:
    return L.make_stats(*ps, pad=pad)


def _def_stats_op(name: str, cfwd: str, cbwd: str):
    @torch.library.custom_op(f"{_NS}::{name}", mutates_args=())
    def fwd(x: Tensor, p1: Tensor, pa: Tensor, pb: Tensor, p3: Tensor, pad: int) -> Tensor:
        _chk(x, p1, pa, pb, p3)
        x = _c(x)
        ps = [_c(p) for p in (p1, pa, pb, p3)]
        out = torch.empty_like(x)
        _call(cfwd, x, _shape5(x), _stats_struct(ps, pad), x, out)
        return out

    @fwd.register_fake
    def _(x, p1, pa, pb, p3, pad):
        return torch.empty_like(x)

    @torch.library.custom_op(f"{_NS}::{name}_bwd", mutates_args=())
    def bwd(x: Tensor, g: Tensor, p1: Tensor, pa: Tensor, pb: Tensor, p3: Tensor, pad: int) -> Tuple[Tensor, Tensor]:
        _chk(x, g, p1, pa, pb, p3)
        x, g = _c(x), _c(g)
        ps = [_c(p) for p in (p1, pa, pb, p3)]
        gx = torch.empty_like(x)
        gst = x.new_zeros(4 * p1.numel())
        _call(cbwd, x, _shape5(x), _stats_struct(ps, pad), x, g, gx, gst)
        return gx, gst

    @bwd.register_fake
    def _(x, g, p1, pa, pb, p3, pad):
        return torch.empty_like(x), x.new_empty(4 * p1.numel())

    def setup(ctx, inputs, output):
        x, p1, pa, pb, p3, pad = inputs
        ctx.save_for_backward(x, p1, pa, pb, p3)
        ctx.pad = pad

    def backward(ctx, g):
        x, p1, pa, pb, p3 = ctx.saved_tensors
        gx, gst = bwd(x, g, p1, pa, pb, p3, ctx.pad)
        n = p1.numel()
        gs = [gst[i * n:(i + 1) * n].reshape(p.shape) for i, p in enumerate((p1, pa, pb, p3))]
        return gx, gs[0], gs[1], gs[2], gs[3], None

    fwd.register_autograd(backward, setup_context=setup)
    return fwd


stats_conv = _def_stats_op("stats_conv", "glrgtv_stats_conv_fwd", "glrgtv_stats_conv_bwd")
stats_conv_t = _def_stats_op("stats_conv_t", "glrgtv_stats_conv_t_fwd", "glrgtv_stats_conv_t_bwd")


# ====================================================================== L, C, Ct
def _out_like_x(x, w, edges):
    return torch.empty_like(x)


def _out_edge(x, w, edges):
    B, G, F, H, W = x.shape
    return x.new_empty(B, G, F, len(edges) // 2, H, W)


def _out_from_edge(z, w, edges):
    B, G, F, E, H, W = z.shape
    return z.new_empty(B, G, F, H, W)


def _def_graph_op(name: str, cfwd: str, cbwd: str, out_fn, shape_from_first: bool):
    """ops of the form out = op(a, w): a is the signal (or the edge signal for Ct), w the edge weights."""

    def shp(a, out):
        ref = a if shape_from_first else out
        return _shape5(ref)

    @torch.library.custom_op(f"{_NS}::{name}", mutates_args=())
    def fwd(a: Tensor, w: Tensor, edges: List[int]) -> Tensor:
        _chk(a, w)
        a, w = _c(a), _c(w)
        out = out_fn(a, w, edges)
        _call(cfwd, a, shp(a, out), _win(edges), a, w, out)
        return out

    fwd.register_fake(out_fn)

    @torch.library.custom_op(f"{_NS}::{name}_bwd", mutates_args=())
    def bwd(a: Tensor, w: Tensor, g: Tensor, edges: List[int]) -> Tuple[Tensor, Tensor]:
        _chk(a, w, g)
        a, w, g = _c(a), _c(w), _c(g)
        ga, gw = torch.empty_like(a), torch.empty_like(w)
        _call(cbwd, a, shp(a, g), _win(edges), a, w, g, ga, gw)
        return ga, gw

    @bwd.register_fake
    def _(a, w, g, edges):
        return torch.empty_like(a), torch.empty_like(w)

    def setup(ctx, inputs, output):
        a, w, edges = inputs
        ctx.save_for_backward(a, w)
        ctx.edges = edges

    def backward(ctx, g):
        a, w = ctx.saved_tensors
        ga, gw = bwd(a, w, g, ctx.edges)
        return ga, gw, None

    fwd.register_autograd(backward, setup_context=setup)
    return fwd


op_L = _def_graph_op("op_L", "glrgtv_op_L_fwd", "glrgtv_op_L_bwd", _out_like_x, True)
op_C = _def_graph_op("op_C", "glrgtv_op_C_fwd", "glrgtv_op_C_bwd", _out_edge, True)
op_Ct = _def_graph_op("op_Ct", "glrgtv_op_Ct_fwd", "glrgtv_op_Ct_bwd", _out_from_edge, False)


# ====================================================================== soft threshold
@torch.library.custom_op(f"{_NS}::soft_threshold", mutates_args=())
def soft_threshold(t: Tensor, thr: Tensor) -> Tensor:
    _chk(t, thr)
    t, thr = _c(t), _c(thr)
    B, G, F, E, H, W = t.shape
    out = torch.empty_like(t)
    _call("glrgtv_soft_threshold_fwd", t, L.make_shape(B, G, F, H, W), E, t, thr, out)
    return out


@soft_threshold.register_fake
def _(t, thr):
    return torch.empty_like(t)


@torch.library.custom_op(f"{_NS}::soft_threshold_bwd", mutates_args=())
def soft_threshold_bwd(t: Tensor, thr: Tensor, g: Tensor) -> Tuple[Tensor, Tensor]:
    _chk(t, thr, g)
    t, thr, g = _c(t), _c(thr), _c(g)
    B, G, F, E, H, W = t.shape
    gt, gthr = torch.empty_like(t), torch.zeros_like(thr)
    _call("glrgtv_soft_threshold_bwd", t, L.make_shape(B, G, F, H, W), E, t, thr, g, gt, gthr)
    return gt, gthr


@soft_threshold_bwd.register_fake
def _(t, thr, g):
    return torch.empty_like(t), torch.empty_like(thr)


def _soft_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _soft_bwd(ctx, g):
    t, thr = ctx.saved_tensors
    return soft_threshold_bwd(t, thr, g)


soft_threshold.register_autograd(_soft_bwd, setup_context=_soft_setup)


# ====================================================================== pooling
def _fine_shape(t: Tensor, fine: bool) -> L.Shape:
    H, W = t.shape[-2:]
    planes = t.numel() // (H * W)
    return L.make_shape(1, 1, planes, H if fine else 2 * H, W if fine else 2 * W)


@torch.library.custom_op(f"{_NS}::pool2", mutates_args=())
def pool2(x: Tensor) -> Tensor:
    _chk(x)
    x = _c(x)
    if x.shape[-1] % 2 or x.shape[-2] % 2:
        raise RuntimeError("pool2 needs even H and W")
    out = x.new_empty(*x.shape[:-2], x.shape[-2] // 2, x.shape[-1] // 2)
    _call("glrgtv_pool2_fwd", x, _fine_shape(x, True), x, out)
    return out


@pool2.register_fake
def _(x):
    return x.new_empty(*x.shape[:-2], x.shape[-2] // 2, x.shape[-1] // 2)


@torch.library.custom_op(f"{_NS}::unpool2", mutates_args=())
def unpool2(x: Tensor) -> Tensor:
    _chk(x)
    x = _c(x)
    out = x.new_empty(*x.shape[:-2], x.shape[-2] * 2, x.shape[-1] * 2)
    _call("glrgtv_unpool2_fwd", x, _fine_shape(x, False), x, out)
    return out


@unpool2.register_fake
def _(x):
    return x.new_empty(*x.shape[:-2], x.shape[-2] * 2, x.shape[-1] * 2)


pool2.register_autograd(lambda ctx, g: unpool2(g))
unpool2.register_autograd(lambda ctx, g: pool2(g))


# ====================================================================== a2 / a3: normalise, gather
@torch.library.custom_op(f"{_NS}::normalize_transform", mutates_args=())
def normalize_transform(feat: Tensor, multiM: Tensor) -> Tensor:
    _chk(feat, multiM)
    feat, multiM = _c(feat), _c(multiM)
    out = torch.empty_like(feat)
    _call("glrgtv_normalize_fwd", feat, _shape5(feat), feat, multiM, out)
    return out


@normalize_transform.register_fake
def _(feat, multiM):
    return torch.empty_like(feat)


@torch.library.custom_op(f"{_NS}::normalize_transform_bwd", mutates_args=())
def normalize_transform_bwd(feat: Tensor, multiM: Tensor, g: Tensor) -> Tuple[Tensor, Tensor]:
    _chk(feat, multiM, g)
    feat, multiM, g = _c(feat), _c(multiM), _c(g)
    gfeat, gM = torch.empty_like(feat), torch.zeros_like(multiM)
    _call("glrgtv_normalize_bwd", feat, _shape5(feat), feat, multiM, g, gfeat, gM)
    return gfeat, gM


@normalize_transform_bwd.register_fake
def _(feat, multiM, g):
    return torch.empty_like(feat), torch.empty_like(multiM)


normalize_transform.register_autograd(
    lambda ctx, g: normalize_transform_bwd(*ctx.saved_tensors, g),
    setup_context=lambda ctx, inputs, output: ctx.save_for_backward(*inputs))


@torch.library.custom_op(f"{_NS}::gather_neighbors", mutates_args=())
def gather_neighbors(x: Tensor, edges: List[int]) -> Tensor:
    _chk(x)
    x = _c(x)
    B, G, F, H, W = x.shape
    out = x.new_empty(B, G, F, len(edges) // 2, H, W)
    _call("glrgtv_gather_neighbors_fwd", x, _shape5(x), _win(edges), x, out)
    return out


@gather_neighbors.register_fake
def _(x, edges):
    B, G, F, H, W = x.shape
    return x.new_empty(B, G, F, len(edges) // 2, H, W)


@torch.library.custom_op(f"{_NS}::gather_neighbors_bwd", mutates_args=())
def gather_neighbors_bwd(g: Tensor, edges: List[int]) -> Tensor:
    _chk(g)
    g = _c(g)
    B, G, F, E, H, W = g.shape
    gx = g.new_empty(B, G, F, H, W)
    _call("glrgtv_gather_neighbors_bwd", g, L.make_shape(B, G, F, H, W), _win(edges), g, gx)
    return gx


@gather_neighbors_bwd.register_fake
def _(g, edges):
    B, G, F, E, H, W = g.shape
    return g.new_empty(B, G, F, H, W)


def _gn_setup(ctx, inputs, output):
    ctx.edges = inputs[1]


gather_neighbors.register_autograd(lambda ctx, g: (gather_neighbors_bwd(g, ctx.edges), None), setup_context=_gn_setup)


# ====================================================================== feature projections (tcgen05 3xTF32 GEMMs, csrc/proj_tc.cu)
def proj_supported(M: int, K: int, N: int) -> bool:
    """shapes glrgtv_proj_gemm / glrgtv_proj_wgrad take: every extent a multiple of 4 (16-byte TMA strides)"""
    return M % 4 == 0 and K % 4 == 0 and N % 4 == 0


@torch.library.custom_op(f"{_NS}::proj_gemm", mutates_args=())
def proj_gemm(w: Tensor, x: Tensor, transpose_w: bool) -> Tensor:
    """w [M,K], x [B,K,N] -> [B,M,N] = w @ x   (transpose_w: x [B,M,N] -> [B,K,N] = w^T @ x); glrgtv_proj_gemm"""
    _chk(w, x)
    w, x = _c(w), _c(x)
    M, K = w.shape
    B, R, N = x.shape
    if R != (M if transpose_w else K):
        raise RuntimeError(f"proj_gemm: weight {tuple(w.shape)} does not match input {tuple(x.shape)}")
    y = x.new_empty(B, K if transpose_w else M, N)
    ws = x.new_empty((int(_lib().glrgtv_proj_gemm_workspace_bytes(M, K)) + 3) // 4)      # the weights as pre-split tile images
    _call("glrgtv_proj_gemm", x, int(bool(transpose_w)), B, M, N, K, w, x, y, ws, ws.numel() * 4)
    return y


@proj_gemm.register_fake
def _(w, x, transpose_w):
    M, K = w.shape
    return x.new_empty(x.shape[0], K if transpose_w else M, x.shape[2])


@torch.library.custom_op(f"{_NS}::proj_wgrad", mutates_args=())
def proj_wgrad(gy: Tensor, x: Tensor) -> Tensor:
    """gy [B,M,N], x [B,K,N] -> gw [M,K] = sum_b gy[b] @ x[b]^T   (glrgtv_proj_wgrad)"""
    _chk(gy, x)
    gy, x = _c(gy), _c(x)
    B, M, N = gy.shape
    K = x.shape[1]
    gw = gy.new_zeros(M, K)
    _call("glrgtv_proj_wgrad", gy, B, M, N, K, gy, x, gw)
    return gw


@proj_wgrad.register_fake
def _(gy, x):
    return gy.new_empty(gy.shape[1], x.shape[1])


def _pg_setup(ctx, inputs, output):
    w, x, transpose_w = inputs
    ctx.save_for_backward(w, x)
    ctx.transpose_w = transpose_w


def _pg_backward(ctx, gy):
    w, x = ctx.saved_tensors
    gy = _c(gy)
    gx = proj_gemm(w, gy, not ctx.transpose_w) if ctx.needs_input_grad[1] else None
    gw = None
    if ctx.needs_input_grad[0]:
        gw = proj_wgrad(x, gy) if ctx.transpose_w else proj_wgrad(gy, x)
    return gw, gx, None


proj_gemm.register_autograd(_pg_backward, setup_context=_pg_setup)


def projection(w: Tensor, x: Tensor) -> Tensor:
    """y[b] = w @ x[b] with gradients, all three GEMMs on the tensor cores"""
    return proj_gemm(w, x, False)


# ---- host CNN, inference forward: the memory-bound pieces of LocalNonLinearBlock (host_cnn.py)
def pixel_rstd(x: Tensor, nsub: int, eps: float) -> Tensor:
    """x [B,C,H,W] -> [B,nsub,H,W]: 1 / sqrt(unbiased variance over each sub-net's channels + eps); glrgtv_pixel_rstd"""
    _chk(x)
    x = _c(x)
    B, C, H, W = x.shape
    rs = x.new_empty(B, nsub, H, W)
    _call("glrgtv_pixel_rstd", x, B, C, nsub, H * W, float(eps), x, rs)
    return rs


def dwconv_gate(h: Tensor, rs: Tensor, w9: Tensor, top: Optional[Tensor] = None, bot: Optional[Tensor] = None) -> Tensor:
    """h [B,2Hd,H,W], rs [B,nsub,H,W], w9 [2Hd,9], top / bot [B,2Hd,W] scaled neighbour rows or None -> u [B,Hd,H,W]
    = sigmoid(g) g v with g | v = the two halves of dw3x3_replicate(rs * h); glrgtv_dwconv_gate"""
    _chk(h, rs, w9, *[t for t in (top, bot) if t is not None])
    h, rs, w9 = _c(h), _c(rs), _c(w9)
    B, C2, H, W = h.shape
    if rs.shape != (B, rs.shape[1], H, W) or w9.shape != (C2, 9) or any(t is not None and t.shape != (B, C2, W) for t in (top, bot)):
        raise RuntimeError("dwconv_gate: inconsistent shapes")
    u = h.new_empty(B, C2 // 2, H, W)
    _call("glrgtv_dwconv_gate", h, B, C2 // 2, rs.shape[1], H, W, h, rs, w9, _c(top) if top is not None else None,
          _c(bot) if bot is not None else None, u)
    return u


def dwconv_gate_bwd(h: Tensor, rs: Tensor, w9: Tensor, gu: Tensor) -> Tuple[Tensor, Tensor]:
    """(gh [B,2Hd,H,W], gw9 [2Hd,9]) from gu = dL/du; glrgtv_dwconv_gate_bwd"""
    _chk(h, rs, w9, gu)
    h, rs, w9, gu = _c(h), _c(rs), _c(w9), _c(gu)
    B, C2, H, W = h.shape
    gM, gh, gw9 = torch.empty_like(h), torch.empty_like(h), torch.zeros_like(w9)
    _call("glrgtv_dwconv_gate_bwd", h, B, C2 // 2, rs.shape[1], H, W, h, rs, w9, gu, gM, gh, gw9)
    return gh, gw9


def pixel_norm_bwd(x: Tensor, rs: Tensor, gx1: Tensor, gout: Tensor, s0: Tensor, nsub: int) -> Tensor:
    """gx = s0 gout + gx1 - <gx1,x>_c rs^2 (x - mean_c x)/(c-1); glrgtv_pixel_norm_bwd"""
    _chk(x, rs, gx1, gout, s0)
    x, rs, gx1, gout, s0 = _c(x), _c(rs), _c(gx1), _c(gout), _c(s0)
    B, C, H, W = x.shape
    gx = torch.empty_like(x)
    _call("glrgtv_pixel_norm_bwd", x, B, C, nsub, H * W, x, rs, gx1, gout, s0, gx)
    return gx


# ---- space-to-depth in front of the 2x2 stride-2 projection (torch.pixel_unshuffle order), and its inverse
@torch.library.custom_op(f"{_NS}::space_to_depth", mutates_args=())
def space_to_depth(x: Tensor, inverse: bool) -> Tensor:
    """x [B,C,H,W] -> [B,4C,H/2,W/2]   (inverse: [B,4C,h,w] -> [B,C,2h,2w]); glrgtv_space_to_depth"""
    _chk(x)
    x = _c(x)
    B, C, H, W = x.shape
    if inverse:
        y = x.new_empty(B, C // 4, 2 * H, 2 * W)
        _call("glrgtv_space_to_depth", x, 1, B * (C // 4), 2 * H, 2 * W, x, y)
    else:
        y = x.new_empty(B, 4 * C, H // 2, W // 2)
        _call("glrgtv_space_to_depth", x, 0, B * C, H, W, x, y)
    return y


@space_to_depth.register_fake
def _(x, inverse):
    B, C, H, W = x.shape
    return x.new_empty(B, C // 4, 2 * H, 2 * W) if inverse else x.new_empty(B, 4 * C, H // 2, W // 2)


def _s2d_setup(ctx, inputs, output):
    ctx.inverse = inputs[1]


space_to_depth.register_autograd(lambda ctx, g: (space_to_depth(g, not ctx.inverse), None), setup_context=_s2d_setup)


# ====================================================================== the fused block (hot path)
# params layout: for each of GTVmodule00, GLRmodule00, GTVmodule01, GLRmodule01: p01, p02a, p02b, p03, multiM (20),
# then alphaCGD, betaCGD, muys00, ro00, gamma00, muys01, ro01, gamma01 (8), then optionally skip_weight (1).
_N_BLOCK_PARAMS = 28
_SAVED = ("wT0", "wL0", "wT1", "wL1", "bA", "x1", "bB", "r1", "x2", "cT0", "cT1", "vc")


def _block_structs(params: Sequence[Tensor]):
    p = L.BlockParams()
    for i, field in enumerate(("gtv0", "glr0", "gtv1", "glr1")):
        op = getattr(p, field)
        op.stats = L.make_stats(*params[5 * i:5 * i + 4])
        op.multiM = params[5 * i + 4].data_ptr()
    for j, field in enumerate(("alpha", "beta", "mu0", "ro0", "gamma0", "mu1", "ro1", "gamma1")):
        setattr(p, field, params[20 + j].data_ptr())
    p.skip = params[28].data_ptr() if len(params) > _N_BLOCK_PARAMS else None
    return p


def _block_geometry(x: Tensor, n_graphs: int):
    B, C, H, W = x.shape
    if C % n_graphs or H % 2 or W % 2:
        raise RuntimeError(f"lowpass_block: C={C} must be a multiple of n_graphs={n_graphs} and H, W even (got {H}x{W})")
    return B, n_graphs, C // n_graphs, H, W


def _saved_shapes(B, G, F, H, W):
    return ([(B, G, 4, H, W)] * 2 + [(B, G, 4, H // 2, W // 2)] * 2 + [(B, G, F, H, W)] * 5 +
            [(B, G, 2, H, W), (B, G, 2, H // 2, W // 2), (2, B, G, F, H // 2, W // 2)])


@torch.library.custom_op(f"{_NS}::lowpass_block_fwd", mutates_args=())
def lowpass_block_fwd(x: Tensor, feat0: Tensor, feat1: Tensor, params: Sequence[Tensor], n_graphs: int) -> List[Tensor]:
    """-> [out, wT0, wL0, wT1, wL1, bA, x1, bB, r1, x2, cT0, cT1]   (glrgtv_block_fwd; 6 kernel launches)"""
    _chk(x, feat0, feat1, *params)
    if len(params) not in (_N_BLOCK_PARAMS, _N_BLOCK_PARAMS + 1):
        raise RuntimeError("lowpass_block: expected 28 (+1 skip) parameter tensors")
    x, feat0, feat1 = _c(x), _c(feat0), _c(feat1)
    params = [_c(p) for p in params]
    B, G, F, H, W = _block_geometry(x, n_graphs)
    out = torch.empty_like(x)
    saved = [x.new_empty(s) for s in _saved_shapes(B, G, F, H, W)]
    sv = L.BlockSaved(*[t.data_ptr() for t in saved])
    _call("glrgtv_block_fwd", x, L.make_shape(B, G, F, H, W), _block_structs(params), x, feat0, feat1, out, sv)
    return [out] + saved


@lowpass_block_fwd.register_fake
def _(x, feat0, feat1, params, n_graphs):
    B, C, H, W = x.shape
    G, F = n_graphs, C // n_graphs
    return [torch.empty_like(x)] + [x.new_empty(s) for s in _saved_shapes(B, G, F, H, W)]


@torch.library.custom_op(f"{_NS}::lowpass_block_bwd", mutates_args=())
def lowpass_block_bwd(x: Tensor, feat0: Tensor, feat1: Tensor, params: Sequence[Tensor], saved: Sequence[Tensor],
                      gout: Tensor, n_graphs: int) -> List[Tensor]:
    """-> [gx, gfeat0, gfeat1, flat parameter gradients]   (glrgtv_block_bwd; 12 kernel launches)"""
    _chk(x, feat0, feat1, gout, *params, *saved)
    x, feat0, feat1, gout = _c(x), _c(feat0), _c(feat1), _c(gout)
    params = [_c(p) for p in params]
    B, G, F, H, W = _block_geometry(x, n_graphs)
    C = G * F
    shp = L.make_shape(B, G, F, H, W)
    has_skip = len(params) > _N_BLOCK_PARAMS
    # one flat zeroed buffer for every parameter gradient (split into per-parameter views by the caller)
    flat = x.new_zeros(sum(_grad_sizes(C, G, F)))
    ptrs, o = [], 0
    for n in _grad_sizes(C, G, F):
        ptrs.append(flat.data_ptr() + 4 * o)
        o += n
    gr = L.BlockGrads(*ptrs)
    if not has_skip:
        gr.skip = None
    nbytes = _lib().glrgtv_block_bwd_workspace_bytes(shp)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=x.device)
    gx, gf0, gf1 = torch.empty_like(x), torch.empty_like(feat0), torch.empty_like(feat1)
    sv = L.BlockSaved(*[t.data_ptr() for t in saved])
    _call("glrgtv_block_bwd", x, shp, _block_structs(params), x, feat0, feat1, sv, gout, gx, gf0, gf1, gr, ws, nbytes)
    return [gx, gf0, gf1, flat]


def _grad_sizes(C, G, F):
    return [4 * C] * 4 + [G * F] * 4 + [3 * G] * 2 + [G] * 6 + [2]


def _split_param_grads(flat: Tensor, params: Sequence[Tensor], G: int, F: int) -> List[Tensor]:
    """flat gradient buffer of glrgtv_block_bwd -> one tensor per entry of the params layout"""
    C = G * F
    views, o = [], 0
    for n in _grad_sizes(C, G, F):
        views.append(flat[o:o + n])
        o += n
    out = []
    for i in range(4):
        out += [views[i][k * C:(k + 1) * C].reshape(params[5 * i + k].shape) for k in range(4)]
        out.append(views[4 + i].reshape(G, F))
    out += [views[8].reshape(3, G), views[9].reshape(3, G)] + [views[10 + j] for j in range(6)]
    if len(params) > _N_BLOCK_PARAMS:
        out.append(views[16])
    return out


@lowpass_block_bwd.register_fake
def _(x, feat0, feat1, params, saved, gout, n_graphs):
    G = n_graphs
    F = x.shape[1] // G
    return [torch.empty_like(x), torch.empty_like(feat0), torch.empty_like(feat1), x.new_empty(sum(_grad_sizes(G * F, G, F)))]


def _lb_setup(ctx, inputs, output):
    x, feat0, feat1, params, n_graphs = inputs
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(x, feat0, feat1, *params, *output[1:])
    ctx.n_params = len(params)
    ctx.n_graphs = n_graphs


def _lb_backward(ctx, grads):
    gout = grads[0]
    t = ctx.saved_tensors
    x, feat0, feat1 = t[:3]
    params, saved = list(t[3:3 + ctx.n_params]), list(t[3 + ctx.n_params:])
    if gout is None:
        return None, None, None, [None] * ctx.n_params, None
    res = lowpass_block_bwd(x, feat0, feat1, params, saved, gout, ctx.n_graphs)
    G = ctx.n_graphs
    return res[0], res[1], res[2], _split_param_grads(res[3], params, G, x.shape[1] // G), None


lowpass_block_fwd.register_autograd(_lb_backward, setup_context=_lb_setup)


def lowpass_block(x: Tensor, feat0: Tensor, feat1: Tensor, params: Sequence[Tensor], n_graphs: int) -> Tensor:
    """LocalLowpassFilteringBlock / MixtureGTVGLR forward (V1X0:707-811, 985-988) as one differentiable op."""
    return lowpass_block_fwd(x, feat0, feat1, list(params), n_graphs)[0]


def lowpass_block_stage(stage: int, x: Tensor, feat0, feat1, params: Sequence[Tensor], n_graphs: int, saved: Sequence[Tensor],
                        out: Tensor, row0: int, row1: int) -> None:
    """One piece of the block forward (glrgtv_block_fwd_stage; inference only, no autograd): stage 0 = edge weights and GTV
    coefficients of the whole plane, 1..4 = BA, X1, X2, X3 on the rows [row0, row1).  `saved` is the list lowpass_block_fwd
    returns after `out` (allocated by the caller: see shard.sharded_block_forward_staged)."""
    _chk(x, out, *params, *saved)
    B, G, F, H, W = _block_geometry(x, n_graphs)
    sv = L.BlockSaved(*[t.data_ptr() for t in saved])
    _call("glrgtv_block_fwd_stage", x, int(stage), L.make_shape(B, G, F, H, W), _block_structs([_c(p) for p in params]), x,
          feat0, feat1, out, sv, int(row0), int(row1))


class PreparedBlockStages:
    """glrgtv_block_fwd_stage with everything marshalled ONCE per plane: the shape, parameter and saved-tensor structs are built (and
    the tensors type-checked) here, a stage call is then a single ctypes call.  A spatially sharded rank makes five stage calls per
    block and image on strips that take well under a millisecond each, so rebuilding ~40 struct fields per call was host time on
    the critical path (tools/strip_host_overhead.py).  Keeps its tensors alive."""

    def __init__(self, x: Tensor, params: Sequence[Tensor], n_graphs: int, saved: Sequence[Tensor], out: Tensor):
        _chk(x, out, *params, *saved)
        B, G, F, H, W = _block_geometry(x, n_graphs)
        self.x, self.out = x, out
        self.params = [_c(p) for p in params]
        self.saved = list(saved)
        self._shape = L.make_shape(B, G, F, H, W)
        self._p = _block_structs(self.params)
        self._sv = L.BlockSaved(*[t.data_ptr() for t in saved])
        self._fn = getattr(_lib(), "glrgtv_block_fwd_stage")
        import ctypes as C
        self._args = (C.byref(self._shape), C.byref(self._p), x.data_ptr())
        self._tail = (out.data_ptr(), C.byref(self._sv))

    def run(self, stage: int, feat0, feat1, row0: int, row1: int) -> None:
        global launch_count
        launch_count += 1
        x = self.x
        with torch.cuda.device(x.device):
            if x.device.index not in _checked_devices:
                L.check(_lib().glrgtv_check_device(), _lib(), f"check_device({x.device})")
                _checked_devices.add(x.device.index)
            rc = self._fn(int(stage), *self._args, None if feat0 is None else feat0.data_ptr(), None if feat1 is None else feat1.data_ptr(),
                          *self._tail, int(row0), int(row1), _stream(x))
        L.check(rc, _lib(), "glrgtv_block_fwd_stage")


def alloc_block_saved(x: Tensor, n_graphs: int) -> List[Tensor]:
    B, G, F, H, W = _block_geometry(x, n_graphs)
    return [x.new_empty(s) for s in _saved_shapes(B, G, F, H, W)]
