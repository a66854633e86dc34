"""Drop-in for the reference module `deep_multiscale_GGLR_GGTV_v1x0` (== model_GLR_GTV_deep_v13 / v22).

Same class names, constructor arguments, attribute names and `state_dict` keys as the reference file
exploration/GGTV_GGLR_v1.0/deep_multiscale_GGLR_GGTV_v1x0.py (V1X0), so that

    import imagerestoration_development_unrolling_b200.deep_multiscale_GGLR_GGTV_v1x0 as model_structure

can replace `import model_GLR_GTV_deep_v13 as model_structure` in the reference's training scripts
(scripts_v2/run_abtract_lightformer_GGTV_GGLR_sigma25.py:33-35, 120-130) and reference checkpoints load
with strict=True.

What differs is what runs: the graph-filter classes (GLRFast, GTVFast, MixtureGTVGLR,
LocalLowpassFilteringBlock; V1X0:13-811, 967-988) call hand-written sm_100a kernels through the C ABI
(ops.py -> libglrgtv.so).  They accept CUDA float32 tensors only; there is no CPU path.  The CNN around
them (V1X0:911-964, 992-1173) is out of the hot path and is plain PyTorch layers with the reference's
parameter layout.
"""
import itertools

import numpy as np
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import ops

_CROSS3 = ((0, 1, 0), (1, 0, 1), (0, 1, 0))
_STATS_INIT = (("stats_kernel_p01", 1.0), ("stats_kernel_p02a", 0.5), ("stats_kernel_p02b", 0.5),
               ("stats_kernel_p03", 0.5))


def _edges_of(mask: np.ndarray):
    """(dh, dw) of the ones of the mask in row-major order (V1X0:42-49)."""
    half = mask.shape[0] // 2
    offs = np.arange(mask.shape[0]) - half
    return [d for d, on in zip(itertools.product(offs, offs), mask.reshape(-1)) if on == 1]


class _GraphOperatorBase(nn.Module):
    """State shared by GLRFast and GTVFast (V1X0:14-125, 243-356): the window, four per-channel
    stats_kernel_p* parameters and multiM."""

    def __init__(self, n_node_fts, n_graphs, M_diag_init=0.4):
        super().__init__()
        self.n_channels = n_node_fts * n_graphs
        self.n_node_fts = n_node_fts
        self.n_graphs = n_graphs
        mask = np.array(_CROSS3)
        self.connection_window = mask
        self.n_edges = int((mask == 1).sum())
        self.buffer_size = int(mask.sum())
        edges = np.array(_edges_of(mask), dtype=np.int32)
        self.edge_delta = torch.tensor(edges, dtype=torch.int32, device="cpu")
        self.pad_dim_hw = torch.tensor(np.abs(edges.min(axis=0)), dtype=torch.int32, device="cpu")
        self._edges_flat = ops.flat_edges(edges.tolist())
        for name, init in _STATS_INIT:
            setattr(self, name, Parameter(torch.full((self.n_channels, 1, 1, 1), init, dtype=torch.float32)))
        self.multiM = Parameter(torch.full((n_graphs, n_node_fts), float(M_diag_init), dtype=torch.float32))

    # -- helpers
    def _stats(self):
        return (self.stats_kernel_p01, self.stats_kernel_p02a, self.stats_kernel_p02b, self.stats_kernel_p03)

    def _as5(self, t):
        if t.dim() == 5:
            return t
        b, _, h, w = t.shape
        return t.reshape(b, self.n_graphs, self.n_node_fts, h, w)

    # -- public methods of the reference
    def get_neighbors_pixels(self, img_features):
        """[B,C,H,W] -> [B,C,E,H,W], replicate-padded neighbours (V1X0:128-144)."""
        b, c, h, w = img_features.shape
        out = ops.gather_neighbors(img_features.reshape(b, 1, c, h, w), self._edges_flat)
        return out.reshape(b, c, self.n_edges, h, w)

    def normalize_and_transform_features(self, img_features):
        """[B,G,F,H,W] -> [B,C,H,W] (V1X0:146-157)."""
        b, g, f, h, w = img_features.shape
        return ops.normalize_transform(img_features, self.multiM).reshape(b, g * f, h, w)

    def extract_edge_weights(self, img_features):
        """[B,G,F,H,W] -> (w [B,G,E,H,W], node_degree [B,G,H,W])  (V1X0:160-175)."""
        w = ops.edge_weights(img_features, self.multiM, self._edges_flat)
        return w, w.sum(dim=2)

    def stats_conv(self, patchs):
        return ops.stats_conv(self._as5(patchs), *self._stats(), 0)

    def stats_conv_transpose(self, patchs):
        return ops.stats_conv_t(self._as5(patchs), *self._stats(), 0)


class GLRFast(_GraphOperatorBase):
    """V1X0:13-237."""

    def op_L_norm(self, img_signals, edge_weights, node_degree):
        return ops.op_L(img_signals, edge_weights, self._edges_flat)

    def forward(self, patchs, edge_weights, node_degree):
        s = self.stats_conv(patchs)
        return self.stats_conv_transpose(self.op_L_norm(s, edge_weights, node_degree))


class GTVFast(_GraphOperatorBase):
    """V1X0:242-523."""

    def op_C(self, img_signals, edge_weights, node_degree):
        return ops.op_C(self.stats_conv(img_signals), edge_weights, self._edges_flat)

    def op_C_transpose(self, edge_signals, edge_weights, node_degree):
        return self.stats_conv_transpose(ops.op_Ct(edge_signals, edge_weights, self._edges_flat))

    def forward(self, patchs, edge_weights, node_degree):
        return self.op_C_transpose(self.op_C(patchs, edge_weights, node_degree), edge_weights, node_degree)


class MixtureGTVGLR(nn.Module):
    """V1X0:526-811.  forward() is ONE fused custom op (ops.lowpass_block) that runs the whole unrolled
    solver in staged sm_100a kernels; the two feature projections are the only library (GEMM) calls."""

    def __init__(self, n_graphs, n_node_fts, alpha_init, beta_init, muy_init, ro_init, gamma_init):
        super().__init__()
        self.n_graphs = n_graphs
        self.n_node_fts = n_node_fts
        self.n_channels = C = n_graphs * n_node_fts
        self.n_cgd_iters = 3
        self.alphaCGD = Parameter(torch.full((3, n_graphs), float(alpha_init), dtype=torch.float32))
        self.betaCGD = Parameter(torch.full((3, n_graphs), float(beta_init), dtype=torch.float32))
        self.scaling_kernel01 = torch.full((C, 1, 2, 2), 0.25, dtype=torch.float32)

        def logp(v):
            return Parameter(torch.ones(n_graphs, dtype=torch.float32) * torch.log(torch.as_tensor(v, dtype=torch.float32).reshape(-1)[0]))

        # registration order == the reference's, so optimizer state in reference checkpoints lines up
        for level in (0, 1):
            sfx = f"0{level}"
            proj = [nn.Conv2d(C, 2 * C, kernel_size=1, bias=False)]
            if level == 1:
                proj.insert(0, nn.Conv2d(C, C, kernel_size=2, stride=2, bias=False))
            setattr(self, "patchs_features_extraction" + sfx, nn.Sequential(*proj))
            setattr(self, "ro" + sfx, logp(ro_init[level]))
            setattr(self, "gamma" + sfx, logp(gamma_init[level]))
            setattr(self, "GTVmodule" + sfx, GTVFast(n_node_fts=n_node_fts, n_graphs=n_graphs, M_diag_init=1.0))
            setattr(self, "muys" + sfx, logp(muy_init[level]))
            setattr(self, "GLRmodule" + sfx, GLRFast(n_node_fts=n_node_fts, n_graphs=n_graphs, M_diag_init=1.0))

    # ---- public helpers of the reference, built from the per-operator kernels
    def soft_threshold(self, delta, gamma):
        return ops.soft_threshold(delta, gamma)

    def apply_lightweight_transformer(self, patchs, graph_weightGTV, graph_weightGLR):
        """A(z) (V1X0:642-682) out of per-operator kernels (the fused path does not call this)."""
        bc = lambda v: torch.exp(v)[None, :, None, None, None]
        z = patchs.contiguous()
        out = z + bc(self.muys00) * self.GLRmodule00(z, *graph_weightGLR[0]) + bc(self.ro00) * self.GTVmodule00(z, *graph_weightGTV[0])
        zc = ops.pool2(z)
        tc = bc(self.muys01) * self.GLRmodule01(zc, *graph_weightGLR[1]) + bc(self.ro01) * self.GTVmodule01(zc, *graph_weightGTV[1])
        return out + ops.unpool2(tc)

    def _block_params(self):
        ps = []
        for m in (self.GTVmodule00, self.GLRmodule00, self.GTVmodule01, self.GLRmodule01):
            ps += [*m._stats(), m.multiM]
        ps += [self.alphaCGD, self.betaCGD, self.muys00, self.ro00, self.gamma00, self.muys01, self.ro01, self.gamma01]
        return ps

    def _projections(self, patchs):
        """patchs_features_extraction00 / 01 (V1X0:556-612, 712, 725) as plain GEMMs on the Conv2d weights: a 1x1 conv is
        W[2C,C] @ x[C,HW]; the 2x2 stride-2 conv is the same after a space-to-depth.  The parameters stay the reference's Conv2d's."""
        b, c, h, w = patchs.shape
        w00 = self.patchs_features_extraction00[0].weight.reshape(2 * c, c)
        w01a = self.patchs_features_extraction01[0].weight.reshape(c, 4 * c)
        w01b = self.patchs_features_extraction01[1].weight.reshape(2 * c, c)
        # all projection GEMMs (forward, input and weight gradients) run on the tcgen05 tensor cores with the three-pass TF32
        # split (csrc/proj_tc.cu): fp32-level accuracy, so the 1e-4 parity bar holds without an fp32 SIMT library GEMM
        def mm(wm, x3):
            if ops.proj_supported(wm.shape[0], wm.shape[1], x3.shape[2]):
                return ops.projection(wm, x3)
            return torch.matmul(wm, x3)              # extents that are not multiples of 4 (no 16-byte TMA rows): library GEMM
        feat0 = mm(w00, patchs.reshape(b, c, h * w)).reshape(b, 2 * c, h, w)
        s2d = ops.space_to_depth(patchs, False) if w % 8 == 0 else nn.functional.pixel_unshuffle(patchs, 2)
        xs = s2d.reshape(b, 4 * c, (h // 2) * (w // 2))
        feat1 = mm(w01b, mm(w01a, xs)).reshape(b, 2 * c, h // 2, w // 2)
        return feat0, feat1

    def forward(self, patchs, _skip_weight=None):
        feat0, feat1 = self._projections(patchs)
        params = self._block_params()
        if _skip_weight is not None:
            params = params + [_skip_weight]
        return ops.lowpass_block(patchs, feat0, feat1, params, self.n_graphs)


# ----------------------------------------------------------------------------------------------------
# CNN around the filter blocks: out of the hot path, plain PyTorch with the reference's parameter layout
# ----------------------------------------------------------------------------------------------------
class CustomLayerNorm(nn.Module):
    """V1X0:911-925: per-pixel variance normalisation within each sub-net, then a per-channel scale."""

    def __init__(self, nchannels, nsubnets):
        super().__init__()
        self.nsubnets, self.nchannels = nsubnets, nchannels
        self.weighted_transform = nn.Conv2d(nchannels, nchannels, kernel_size=1, groups=nchannels, bias=False)

    def forward(self, x):
        b, c, h, w = x.shape
        v = x.reshape(b, self.nsubnets, c // self.nsubnets, h, w)
        v = v / torch.sqrt(v.var(dim=2, keepdim=True, correction=1) + 1e-5)
        return self.weighted_transform(v.reshape(b, c, h, w))


class LocalGatedLinearBlock(nn.Module):
    """V1X0:929-948."""

    def __init__(self, dim, hidden_dim, nsubnets):
        super().__init__()
        self.channels_linear_op = nn.Conv2d(dim, 2 * hidden_dim, kernel_size=1, bias=False, groups=nsubnets)
        self.channels_local_linear_op = nn.Conv2d(2 * hidden_dim, 2 * hidden_dim, kernel_size=3, padding=1,
                                                  padding_mode="replicate", groups=2 * hidden_dim, bias=False)
        self.project_out = nn.Conv2d(hidden_dim, dim, kernel_size=1, bias=False, groups=nsubnets)

    def forward(self, x):
        gate, val = self.channels_local_linear_op(self.channels_linear_op(x)).chunk(2, dim=1)
        return self.project_out(torch.sigmoid(gate) * gate * val)


HOST_CNN_KERNELS = True


def set_host_cnn_kernels(on: bool) -> bool:
    """Every LocalNonLinearBlock of the host CNN on libglrgtv's kernels (host_cnn.py: forward under no_grad, forward + backward
    under autograd; the 1x1 convolutions on the tcgen05 GEMM) instead of the PyTorch op sequence, for CUDA fp32 inputs whose width
    is a multiple of 4.  Same results (tests/test_gpu_host_cnn*.py, tests/test_gpu_zz_model_switch.py).  ON by default since it was
    timed (profiles/r02_configs.md: the whole v13 network's training step at 4 x 256^2 424 -> 109 ms, peak memory 61 -> 21 GB);
    False restores the module's own PyTorch layers.  Returns the previous setting.  Not used while torch.compile is tracing."""
    global HOST_CNN_KERNELS
    prev, HOST_CNN_KERNELS = HOST_CNN_KERNELS, bool(on)
    return prev


FILTER_STREAMS = True
_side_streams = {}


def set_filter_streams(on: bool) -> bool:
    """`AbtractMultiScaleGraphFilter.filtering` runs its four LocalLowpassFilteringBlocks on four CUDA streams (the caller's
    stream + three side streams, fork / join by events) - the blocks are independent (V1X0:1117-1131).  The same kernels and
    results; it shortens the latency of small batches, where one block cannot fill 148 SMs (config 1: 1.19 -> 0.61 ms inside a
    CUDA graph), overlaps the tails of large ones (the bench's training step: 31.3 -> 28.5 ms) and is capturable in a CUDA graph.
    Autograd runs each block's backward on the stream its forward ran on.  ON by default for CUDA inputs (not while
    torch.compile is tracing); False runs the blocks one after the other.  Returns the previous setting."""
    global FILTER_STREAMS
    prev, FILTER_STREAMS = FILTER_STREAMS, bool(on)
    return prev


def run_blocks_on_streams(blocks, inputs):
    """outs[i] = blocks[i](inputs[i]); block 0 on the current stream, the others on per-device side streams that fork from and
    join back into it."""
    cur = torch.cuda.current_stream(inputs[0].device)
    key = (inputs[0].device.index, len(blocks) - 1)
    if key not in _side_streams:
        _side_streams[key] = [torch.cuda.Stream(inputs[0].device) for _ in range(len(blocks) - 1)]
    side = _side_streams[key]
    fork = torch.cuda.Event()
    fork.record(cur)
    outs = [None] * len(blocks)
    for i in range(1, len(blocks)):
        side[i - 1].wait_event(fork)
        with torch.cuda.stream(side[i - 1]):
            outs[i] = blocks[i](inputs[i])
        inputs[i].record_stream(side[i - 1])              # the caching allocator must not recycle them under the side stream
    outs[0] = blocks[0](inputs[0])
    for i in range(1, len(blocks)):
        cur.wait_stream(side[i - 1])
        outs[i].record_stream(cur)
    return outs


class LocalNonLinearBlock(nn.Module):
    """V1X0:951-964."""

    def __init__(self, dim, hidden_dim, nsubnets):
        super().__init__()
        self.norm = CustomLayerNorm(dim, nsubnets)
        self.local_linear = LocalGatedLinearBlock(dim, hidden_dim, nsubnets)
        self.skip_weight = Parameter(torch.tensor([1.0, 1.0], dtype=torch.float32))

    def forward(self, x):
        if (HOST_CNN_KERNELS and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[-1] % 4 == 0
                and not torch.compiler.is_compiling()):
            from . import host_cnn
            if torch.is_grad_enabled():
                return host_cnn.nonlinear_block_train(self, x)
            return host_cnn.nonlinear_block_forward(self, x)
        return self.skip_weight[0] * x + self.skip_weight[1] * self.local_linear(self.norm(x))


class LocalLowpassFilteringBlock(nn.Module):
    """V1X0:967-988.  `skip_weight` is folded into the last fused kernel."""

    def __init__(self, dim, nsubnets, ngraphs):
        super().__init__()
        self.local_filter = MixtureGTVGLR(
            n_graphs=ngraphs, n_node_fts=dim // ngraphs, alpha_init=0.5, beta_init=0.1,
            muy_init=torch.tensor([[0.001], [0.0001]]), ro_init=torch.tensor([[0.0001], [0.0001]]),
            gamma_init=torch.tensor([[0.0001], [0.0001]]))
        self.skip_weight = Parameter(torch.tensor([0.5, 0.5], dtype=torch.float32))

    def forward(self, x):
        return self.local_filter(x, _skip_weight=self.skip_weight)


class ReginalPixelEmbeding(nn.Module):
    """V1X0:992-1005."""

    def __init__(self, n_channels_in=3, dim=48, bias=False):
        super().__init__()
        self.channels_local_linear_op01 = nn.Conv2d(n_channels_in, dim, kernel_size=3, padding=1,
                                                    padding_mode="replicate", bias=False)

    def forward(self, x):
        return self.channels_local_linear_op01(x)


class Downsampling(nn.Module):
    """V1X0:1010-1016."""

    def __init__(self, dim_in, dim_out, nsubnets):
        super().__init__()
        self.local_linear = nn.Conv2d(dim_in, dim_out, kernel_size=2, stride=2, groups=nsubnets, bias=False)

    def forward(self, x):
        return self.local_linear(x)


class Upsampling(nn.Module):
    """V1X0:1018-1024."""

    def __init__(self, dim_in, dim_out, nsubnets):
        super().__init__()
        self.local_linear = nn.ConvTranspose2d(dim_in, dim_out, kernel_size=2, stride=2, groups=nsubnets, bias=False)

    def forward(self, x):
        return self.local_linear(x)


class AbtractMultiScaleGraphFilter(nn.Module):
    """V1X0:1028-1173: 4-scale encoder -> one filter block per scale -> decoder."""

    def __init__(self, n_channels_in=3, n_channels_out=3, dims=[48, 64, 96, 128], hidden_dims=[128, 192, 256, 384],
                 nsubnets=[1, 1, 1, 1], ngraphs=[4, 4, 8, 8], num_blocks=[4, 6, 6, 8], num_blocks_out=4):
        super().__init__()

        def stack(i, n):
            return nn.Sequential(*[LocalNonLinearBlock(dims[i], hidden_dims[i], nsubnets[i]) for _ in range(n)])

        self.patch_3x3_embeding = ReginalPixelEmbeding(n_channels_in, dims[0])
        self.encoder_scale_00 = stack(0, num_blocks[0])
        for i in (1, 2, 3):
            setattr(self, f"down_sample_0{i - 1}_0{i}", Downsampling(dims[i - 1], dims[i], nsubnets[i - 1]))
            setattr(self, f"encoder_scale_0{i}", stack(i, num_blocks[i]))
        for i in range(4):
            setattr(self, f"localfilter_scale_0{i}", LocalLowpassFilteringBlock(dims[i], nsubnets[i], ngraphs[i]))
        for i in (2, 1, 0):
            setattr(self, f"up_sample_0{i + 1}_0{i}", Upsampling(dims[i + 1], dims[i], nsubnets[i + 1]))
            setattr(self, f"combine_channels_0{i}", nn.Conv2d(2 * dims[i], dims[i], kernel_size=1, bias=False, groups=nsubnets[i]))
            setattr(self, f"decoder_scale_0{i}", stack(i, num_blocks[i]))
        self.refining_block = stack(0, num_blocks_out)
        self.linear_output = nn.Conv2d(dims[0], n_channels_out, kernel_size=1, bias=False)

    def encode(self, img):
        x = self.encoder_scale_00(self.patch_3x3_embeding(img))
        outs = [x]
        for i in (1, 2, 3):
            x = getattr(self, f"encoder_scale_0{i}")(getattr(self, f"down_sample_0{i - 1}_0{i}")(x))
            outs.append(x)
        return tuple(outs)

    def filtering(self, coefs):
        blocks = [getattr(self, f"localfilter_scale_0{i}") for i in range(len(coefs))]
        if FILTER_STREAMS and coefs[0].is_cuda and not torch.compiler.is_compiling():
            return tuple(run_blocks_on_streams(blocks, list(coefs)))
        return tuple(b(c) for b, c in zip(blocks, coefs))

    def decode(self, coefs):
        x = coefs[3]
        for i in (2, 1, 0):
            up = getattr(self, f"up_sample_0{i + 1}_0{i}")(x)
            x = getattr(self, f"combine_channels_0{i}")(torch.cat([up, coefs[i]], 1))
            x = getattr(self, f"decoder_scale_0{i}")(x)
        return self.linear_output(self.refining_block(x))

    def enc_dec(self, img):
        return self.decode(self.encode(img))

    def forward(self, img):
        return self.decode(self.filtering(self.encode(img)))
