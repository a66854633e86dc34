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
from typing import List, Sequence, Tuple

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


def _call(name: str, ref: Tensor, *args):
    global launch_count
    launch_count += 1
    with torch.cuda.device(ref.device):
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
