"""Inference forward of the host CNN's LocalNonLinearBlock (V1X0:911-964) on the library's kernels (SURVEY 8f rank 1,
the HBM-bound pieces on own kernels; the two 1x1 convolutions on the tcgen05 three-pass-TF32 GEMM of csrc/proj_tc.cu, see GEMM).

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
# the two 1x1 convolutions of the inference forward: "tc" = libglrgtv's tcgen05 three-pass-TF32 GEMM (csrc/proj_tc.cu, the
# projection kernel of the filter blocks; fp32-level accuracy), "cublas" = torch.matmul (fp32 SIMT with TF32 off); "auto" = "tc"
# on CUDA tensors whose extents the kernel takes (multiples of 4, one sub-network), else the library GEMM
GEMM = "auto"


def _mm(w: torch.Tensor, x3: torch.Tensor) -> torch.Tensor:
    """w [nsub, M, K], x3 [B, nsub, K, N] -> [B, nsub, M, N]"""
    if GEMM != "cublas" and x3.is_cuda and w.shape[0] == 1:
        from . import ops
        if ops.proj_supported(w.shape[1], w.shape[2], x3.shape[-1]):
            return ops.proj_gemm(w[0], x3[:, 0], False).unsqueeze(1)
    return torch.matmul(w, x3)


class CudaCnnKernels:
    def pixel_rstd(self, x: torch.Tensor, nsub: int, eps: float) -> torch.Tensor:
        from . import ops
        return ops.pixel_rstd(x, nsub, eps)

    def dwconv_gate(self, h, rs, w9, top, bot) -> torch.Tensor:
        from . import ops
        return ops.dwconv_gate(h, rs, w9, top, bot)

    def dwconv_gate_bwd(self, h, rs, w9, gu):
        from . import ops
        return ops.dwconv_gate_bwd(h, rs, w9, gu)

    def pixel_norm_bwd(self, x, rs, gx1, gout, s0, nsub):
        from . import ops
        return ops.pixel_norm_bwd(x, rs, gx1, gout, s0, nsub)


def _mm_t(w: torch.Tensor, g4: torch.Tensor) -> torch.Tensor:
    """w [nsub, M, K], g4 [B, nsub, M, N] -> [B, nsub, K, N] = w^T g (input gradient of _mm)"""
    if GEMM != "cublas" and g4.is_cuda and w.shape[0] == 1:
        from . import ops
        if ops.proj_supported(w.shape[1], w.shape[2], g4.shape[-1]):
            return ops.proj_gemm(w[0].contiguous(), g4[:, 0], True).unsqueeze(1)
    return torch.matmul(w.transpose(-1, -2), g4)


def _wgrad(g4: torch.Tensor, x4: torch.Tensor) -> torch.Tensor:
    """g4 [B, nsub, M, N], x4 [B, nsub, K, N] -> [nsub, M, K] = sum_b g x^T (weight gradient of _mm)"""
    if GEMM != "cublas" and g4.is_cuda and g4.shape[1] == 1:
        from . import ops
        if ops.proj_supported(g4.shape[2], x4.shape[2], g4.shape[-1]):
            return ops.proj_wgrad(g4[:, 0], x4[:, 0]).unsqueeze(0)
    return torch.matmul(g4, x4.transpose(-1, -2)).sum(0)


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
    h = _mm(w1, x.view(B, nsub, C // nsub, H * W)).view(B, -1, H, W)                    # [B, 2Hd, H, W], un-normalised
    top = bot = None
    if exchange is not None:
        per = h.shape[1] // nsub
        scaled = lambda r: (h[:, :, r].view(B, nsub, per, W) * rs[:, :, r].unsqueeze(2)).reshape(B, -1, W).contiguous()  # noqa: E731
        top, bot = exchange(scaled(0), scaled(H - 1))
    u = kernels.dwconv_gate(h, rs, w9, top, bot)                                        # [B, Hd, H, W]
    y = _mm(w2, u.view(B, nsub, -1, H * W)).view(B, C, H, W)
    return torch.addcmul(y, x, s0)


# ---------------------------------------------------------------------------------------------------- training
class _NonLinearBlockFn(torch.autograd.Function):
    """LocalNonLinearBlock forward + backward on the kernels.  Saves x, rs, h (un-normalised 1x1 output) and u - the module's
    autograd keeps the normalised input, the padded hidden tensor, both halves, the sigmoid and two products besides."""

    @staticmethod
    def forward(ctx, x, w_norm, w_lin, w_dw, w_out, skip, nsub, kernels):
        B, C, H, W = x.shape
        x = x.contiguous()
        w1 = (w_lin.reshape(nsub, w_lin.shape[0] // nsub, w_lin.shape[1]) * w_norm.reshape(nsub, 1, -1)).contiguous()
        w2 = w_out.reshape(nsub, w_out.shape[0] // nsub, w_out.shape[1])
        w9 = w_dw.reshape(-1, 9).contiguous()
        rs = kernels.pixel_rstd(x, nsub, NORM_EPS)
        h = _mm(w1, x.view(B, nsub, C // nsub, H * W)).view(B, -1, H, W)
        u = kernels.dwconv_gate(h, rs, w9, None, None)
        y = _mm((w2 * skip[1]).contiguous(), u.view(B, nsub, -1, H * W)).view(B, C, H, W)
        ctx.save_for_backward(x, rs, h, u, w_norm, w_lin, w_dw, w_out, skip)
        ctx.nsub, ctx.kernels = nsub, kernels
        return torch.addcmul(y, x, skip[0:1])

    @staticmethod
    def backward(ctx, gout):
        x, rs, h, u, w_norm, w_lin, w_dw, w_out, skip = ctx.saved_tensors
        nsub, kernels = ctx.nsub, ctx.kernels
        B, C, H, W = x.shape
        N = H * W
        gout = gout.contiguous()
        w1 = w_lin.reshape(nsub, w_lin.shape[0] // nsub, w_lin.shape[1])
        wn = w_norm.reshape(nsub, 1, -1)
        w1f = (w1 * wn).contiguous()
        w2 = w_out.reshape(nsub, w_out.shape[0] // nsub, w_out.shape[1])
        g4, u4, x4 = gout.view(B, nsub, C // nsub, N), u.view(B, nsub, -1, N), x.view(B, nsub, C // nsub, N)
        # out = s0 x + s1 W2 u
        gw2_raw = _wgrad(g4, u4)                                                             # [nsub, C/nsub, Hd/nsub] = sum gout u^T
        g_skip = torch.stack([torch.dot(gout.reshape(-1), x.reshape(-1)), (gw2_raw * w2).sum()])
        gu = _mm_t((w2 * skip[1]).contiguous(), g4).view(B, -1, H, W)                        # [B, Hd, H, W]
        gh, gw9 = kernels.dwconv_gate_bwd(h, rs, w_dw.reshape(-1, 9).contiguous(), gu)
        gh4 = gh.view(B, nsub, -1, N)
        gw1f = _wgrad(gh4, x4)                                                               # [nsub, 2Hd/nsub, C/nsub]
        gx1 = _mm_t(w1f, gh4).view(B, C, H, W)
        gx = kernels.pixel_norm_bwd(x, rs, gx1, gout, skip[0:1], nsub)
        return (gx, (gw1f * w1).sum(1).reshape(w_norm.shape), (gw1f * wn).reshape(w_lin.shape), gw9.reshape(w_dw.shape),
                (gw2_raw * skip[1]).reshape(w_out.shape), g_skip, None, None)


def nonlinear_block_train(blk, x: torch.Tensor, kernels=None) -> torch.Tensor:
    """LocalNonLinearBlock(x) with autograd through the kernels (opt-in: the drop-in module's own forward stays PyTorch until this
    path has been measured on the GPU).  Whole images (training patches), W % 4 == 0."""
    ll = blk.local_linear
    return _NonLinearBlockFn.apply(x, blk.norm.weighted_transform.weight, ll.channels_linear_op.weight.flatten(1),
                                   ll.channels_local_linear_op.weight, ll.project_out.weight.flatten(1), blk.skip_weight,
                                   blk.norm.nsubnets, kernels or CudaCnnKernels())
