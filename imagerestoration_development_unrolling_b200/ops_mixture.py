"""Mixture weighting of the older model family as a custom op (V7 = model_GLR_GTV_deep_v7.py:847-858, 1011-1014):
out[b,c,h,w] = sum_g x[b,g,c,h,w] * score[b,g,h,w].  CUDA only, through glrgtv_mixture_fwd / _bwd."""
from typing import Tuple

import torch
from torch import Tensor

from . import _lib as L
from .ops import _NS, _c, _call, _chk


@torch.library.custom_op(f"{_NS}::mixture", mutates_args=())
def mixture(x: Tensor, score: Tensor) -> Tensor:
    _chk(x, score)
    x, score = _c(x), _c(score)
    B, G, F, H, W = x.shape
    out = x.new_empty(B, F, H, W)
    _call("glrgtv_mixture_fwd", x, L.make_shape(B, G, F, H, W), x, score, out)
    return out


@mixture.register_fake
def _(x, score):
    B, G, F, H, W = x.shape
    return x.new_empty(B, F, H, W)


@torch.library.custom_op(f"{_NS}::mixture_bwd", mutates_args=())
def mixture_bwd(x: Tensor, score: Tensor, gout: Tensor) -> Tuple[Tensor, Tensor]:
    _chk(x, score, gout)
    x, score, gout = _c(x), _c(score), _c(gout)
    B, G, F, H, W = x.shape
    gx, gs = torch.empty_like(x), torch.empty_like(score)
    _call("glrgtv_mixture_bwd", x, L.make_shape(B, G, F, H, W), x, score, gout, gx, gs)
    return gx, gs


@mixture_bwd.register_fake
def _(x, score, gout):
    return torch.empty_like(x), torch.empty_like(score)


mixture.register_autograd(lambda ctx, g: mixture_bwd(*ctx.saved_tensors, g),
                          setup_context=lambda ctx, inputs, output: ctx.save_for_backward(*inputs))
