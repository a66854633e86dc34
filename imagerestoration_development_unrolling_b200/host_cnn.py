"""Inference forward of the host CNN's LocalNonLinearBlock (V1X0:911-964) on the library's kernels (SURVEY 8f rank 1,
first cut: the HBM-bound part; the two 1x1 convolutions stay cuBLAS fp32 GEMMs).

    reference                                   here
    n = w_n * x / sqrt(var_c(x) + 1e-5)         rs = glrgtv_pixel_rstd(x)                       (x read once)
    h = conv1x1(n)                              h  = (W1 diag(w_n)) @ x                          (rs commutes with the 1x1)
    m = dw3x3_replicate(h); g, v = m.chunk(2)   u  = glrgtv_dwconv_gate(h, rs, w_dw, top, bot)   (h read once, u written)
    out = s0 x + s1 conv1x1(sigmoid(g) g v)     out = addcmul((s1 W2) @ u, x, s0)

`kernels` supplies the two kernel calls (CudaCnnKernels: libglrgtv.so on CUDA tensors; the CPU tests substitute the
emulation build).  `exchange(first_row, last_row) -> (top, bot)` supplies the neighbours' rows on a row strip of a
spatially sharded image (shard.ShardedMultiScaleFilter); None = whole image.  No autograd: training keeps the module."""
from typing import Callable, Optional, Tuple

import torch

NORM_EPS = 1e-5      # V1X0:921


class CudaCnnKernels:
    def pixel_rstd(self, x: torch.Tensor, nsub: int, eps: float) -> torch.Tensor:
        from . import ops
        return ops.pixel_rstd(x, nsub, eps)

    def dwconv_gate(self, h, rs, w9, top, bot) -> torch.Tensor:
        from . import ops
        return ops.dwconv_gate(h, rs, w9, top, bot)


def folded_weights(blk) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(W1' [nsub, 2Hd/nsub, C/nsub], w_dw [2Hd, 9], s1 W2 [nsub, C/nsub, Hd/nsub], s0 [1]) of a LocalNonLinearBlock"""
    nsub = blk.norm.nsubnets
    ll = blk.local_linear
    w_n = blk.norm.weighted_transform.weight.reshape(nsub, 1, -1)                       # [nsub, 1, C/nsub]
    w1 = ll.channels_linear_op.weight
    w1 = w1.reshape(nsub, w1.shape[0] // nsub, w1.shape[1]) * w_n
    w2 = ll.project_out.weight
    w2 = w2.reshape(nsub, w2.shape[0] // nsub, w2.shape[1]) * blk.skip_weight[1]
    w9 = ll.channels_local_linear_op.weight.reshape(-1, 9).contiguous()
    return w1.contiguous(), w9, w2.contiguous(), blk.skip_weight[0:1]


@torch.no_grad()
def nonlinear_block_forward(blk, x: torch.Tensor, kernels=None,
                            exchange: Optional[Callable[[torch.Tensor, torch.Tensor], Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]]] = None,
                            ) -> torch.Tensor:
    kernels = kernels or CudaCnnKernels()
    if blk.local_linear.channels_local_linear_op.padding_mode != "replicate":
        raise ValueError("LocalNonLinearBlock's depthwise convolution is replicate padded (V1X0:938-943)")
    B, C, H, W = x.shape
    nsub = blk.norm.nsubnets
    w1, w9, w2, s0 = folded_weights(blk)
    x = x.contiguous()
    rs = kernels.pixel_rstd(x, nsub, NORM_EPS)                                          # [B, nsub, H, W]
    h = torch.matmul(w1, x.view(B, nsub, C // nsub, H * W)).view(B, -1, H, W)           # [B, 2Hd, H, W], un-normalised
    top = bot = None
    if exchange is not None:
        per = h.shape[1] // nsub
        scaled = lambda r: (h[:, :, r].view(B, nsub, per, W) * rs[:, :, r].unsqueeze(2)).reshape(B, -1, W).contiguous()  # noqa: E731
        top, bot = exchange(scaled(0), scaled(H - 1))
    u = kernels.dwconv_gate(h, rs, w9, top, bot)                                        # [B, Hd, H, W]
    y = torch.matmul(w2, u.view(B, nsub, -1, H * W)).view(B, C, H, W)
    return torch.addcmul(y, x, s0)
