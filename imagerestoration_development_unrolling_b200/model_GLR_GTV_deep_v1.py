"""Drop-in for the reference's three-block chain, exploration/model_multiscale_mixture_GLR/lib/model_GLR_GTV_deep_v1.py (V1;
SURVEY 8a row a20: "v1: three blocks (3x3 full / 8 edges x2, 5x5 full / 24 edges), 6 iterations each, SharpeningBlock
between").

Same class names, constructor arguments and `state_dict` keys as V1.  What differs from the V7 family (model_GLR_GTV_deep_v7.py
in this package): the graph operators carry no stats convolution (their only parameter is multiM), the solver schedule is
2 + 4 momentum iterations around ONE threshold / dual update (V1:602-676), the feature CNN is a four-level U-Net and there is no
DC estimator.  The graph operators and the unrolled solver run libglrgtv's kernels through ops.py (generic-window family-A
operators); the feature CNN and the SharpeningBlocks are out of the hot path and are plain PyTorch layers with V1's layout.

  FeatureExtraction              V1:108-184
  GLRFast / GTVFast              V1:187-291, 293-470
  MixtureGTV                     V1:472-676
  SharpeningBlock                V1:768-787
  MultiScaleSequenceDenoiser     V1:790-883
"""
import itertools

import numpy as np
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import model_GLR_GTV_deep_v7 as _v7
from . import ops
from .model_GLR_GTV_deep_v7 import (CustomLayerNorm, Downsample, FeedForward, FFBlock, OverlapPatchEmbed,  # noqa: F401  (V1:13-106,
                                    Upsample)                                                              # same layers as V7's)
from .ops_mixture import mixture


class FeatureExtraction(nn.Module):
    """four-level conv U-Net (V1:108-184); returns [level-1 features, level-2, level-3, latent] like the reference"""

    def __init__(self, inp_channels=3, out_channels=48, dim=48, num_blocks=[1, 2, 2, 4], num_refinement_blocks=4,
                 ffn_expansion_factor=2.66, bias=False):
        super().__init__()
        blocks = lambda d, n: nn.Sequential(*[FFBlock(d, ffn_expansion_factor, bias) for _ in range(n)])
        self.patch_embed = OverlapPatchEmbed(inp_channels, dim)
        self.encoder_level1 = blocks(dim, num_blocks[0])
        self.down1_2 = Downsample(dim)
        self.encoder_level2 = blocks(2 * dim, num_blocks[1])
        self.down2_3 = Downsample(2 * dim)
        self.encoder_level3 = blocks(4 * dim, num_blocks[2])
        self.down3_4 = Downsample(4 * dim)
        self.latent = blocks(8 * dim, num_blocks[3])
        self.up4_3 = Upsample(8 * dim)
        self.reduce_chan_level3 = nn.Conv2d(8 * dim, 4 * dim, kernel_size=1, bias=bias)
        self.decoder_level3 = blocks(4 * dim, num_blocks[2])
        self.up3_2 = Upsample(4 * dim)
        self.reduce_chan_level2 = nn.Conv2d(4 * dim, 2 * dim, kernel_size=1, bias=bias)
        self.decoder_level2 = blocks(2 * dim, num_blocks[1])
        self.up2_1 = Upsample(2 * dim)
        self.decoder_level1 = blocks(2 * dim, num_blocks[0])
        self.refinement = blocks(2 * dim, num_refinement_blocks)
        self.output = nn.Conv2d(2 * dim, out_channels, kernel_size=3, padding=1, bias=bias)

    def forward(self, inp_img):
        e1 = self.encoder_level1(self.patch_embed(inp_img))
        e2 = self.encoder_level2(self.down1_2(e1))
        e3 = self.encoder_level3(self.down2_3(e2))
        latent = self.latent(self.down3_4(e3))
        d3 = self.decoder_level3(self.reduce_chan_level3(torch.cat([self.up4_3(latent), e3], 1)))
        d2 = self.decoder_level2(self.reduce_chan_level2(torch.cat([self.up3_2(d3), e2], 1)))
        d1 = self.refinement(self.decoder_level1(torch.cat([self.up2_1(d2), e1], 1)))
        return [self.output(d1), d2, d3, latent]


class _GraphOperatorV1(nn.Module):
    """V1:187-272 / 293-419: window from a 0/1 mask, replicate-padded neighbours, multiM [G,F]; no stats convolution"""

    def __init__(self, n_channels, n_node_fts, n_graphs, connection_window, device, M_diag_init=0.4):
        super().__init__()
        self.device = device
        self.n_channels, self.n_node_fts, self.n_graphs = n_channels, n_node_fts, n_graphs
        mask = np.asarray(connection_window)
        self.n_edges = int((mask == 1).sum())
        self.connection_window = mask
        self.buffer_size = int(mask.sum())
        offs = np.arange(mask.shape[0]) - mask.shape[0] // 2
        self.edge_delta = np.array([d for d, on in zip(itertools.product(offs, offs), mask.reshape(-1)) if on == 1], dtype=np.int32)
        self.pad_dim_hw = np.abs(self.edge_delta.min(axis=0))
        self._edges_flat = ops.flat_edges(self.edge_delta.tolist())
        self.multiM = Parameter(torch.full((n_graphs, n_node_fts), float(M_diag_init), dtype=torch.float32, device=device))

    get_neighbors_pixels = _v7._GraphOperatorA.get_neighbors_pixels
    normalize_and_transform_features = _v7._GraphOperatorA.normalize_and_transform_features
    extract_edge_weights = _v7._GraphOperatorA.extract_edge_weights
    _over_graphs = _v7._GraphOperatorA._over_graphs


class GLRFast(_GraphOperatorV1):
    def op_L_norm(self, img_signals, edge_weights, node_degree):
        return ops.op_L(self._over_graphs(img_signals, edge_weights), edge_weights, self._edges_flat)

    def forward(self, patchs, edge_weights, node_degree):
        return self.op_L_norm(patchs, edge_weights, node_degree)


class GTVFast(_GraphOperatorV1):
    def op_C(self, img_signals, edge_weights, node_degree):
        return ops.op_C(self._over_graphs(img_signals, edge_weights), edge_weights, self._edges_flat)

    def op_C_transpose(self, edge_signals, edge_weights, node_degree):
        return ops.op_Ct(edge_signals.contiguous(), edge_weights, self._edges_flat)

    def forward(self, patchs, edge_weights, node_degree):
        return self.op_C_transpose(self.op_C(patchs, edge_weights, node_degree), edge_weights, node_degree)


class MixtureGTV(_v7.MixtureGTV):
    """V1:472-676: 2 momentum iterations, one threshold with the dual update, 4 more iterations, mixture output.  The solver loop
    is the V7 class's `unrolled_solve` with schedule (2, 4)."""

    def __init__(self, nchannels_in, n_graphs, n_node_fts, connection_window, n_cgd_iters, alpha_init, beta_init, muy_init, ro_init,
                 gamma_init, device):
        nn.Module.__init__(self)
        self.device = device
        self.n_graphs, self.n_node_fts = n_graphs, n_node_fts
        self.n_total_fts = n_graphs * n_node_fts
        self.n_levels, self.n_cgd_iters = 4, n_cgd_iters
        self.nchannels_in, self.connection_window = nchannels_in, connection_window
        vec = lambda v: Parameter(torch.ones(n_graphs, dtype=torch.float32, device=device) * v)
        self.alphaCGD = Parameter(torch.full((n_cgd_iters, n_graphs), float(alpha_init), dtype=torch.float32, device=device))
        self.betaCGD = Parameter(torch.full((n_cgd_iters, n_graphs), float(beta_init), dtype=torch.float32, device=device))
        self.patchs_features_extraction = FeatureExtraction(
            inp_channels=3, out_channels=self.n_total_fts, dim=self.n_total_fts, num_blocks=[2, 2, 2, 2], num_refinement_blocks=4,
            ffn_expansion_factor=1, bias=False).to(device)
        self.combination_weight = nn.Sequential(nn.Conv2d(self.n_total_fts, n_graphs, kernel_size=1, bias=False),
                                                nn.Softmax(dim=1)).to(device)
        scalar = lambda t: torch.as_tensor(t, dtype=torch.float32).reshape(-1)[0].to(device)
        self.ro00 = vec(scalar(ro_init[0]))                           # raw (V1:529-532)
        self.gamma00 = vec(torch.log(scalar(gamma_init[0])))          # log domain (V1:533-536)
        self.GTVmodule00 = GTVFast(nchannels_in, n_node_fts, n_graphs, connection_window, device, M_diag_init=1.0)
        self.muys00 = vec(scalar(muy_init[0]))
        self.GLRmodule00 = GLRFast(nchannels_in, n_node_fts, n_graphs, connection_window, device, M_diag_init=1.0)

    def forward(self, patchs):
        feats = self.patchs_features_extraction(patchs)[0]
        b, _, h, w = feats.shape
        gfeat = feats.reshape(b, self.n_graphs, self.n_node_fts, h, w)
        wT, wL = self.GTVmodule00.extract_edge_weights(gfeat), self.GLRmodule00.extract_edge_weights(gfeat)
        out = self.unrolled_solve(patchs[:, None], wT, wL, schedule=(2, 4))
        return mixture(out.contiguous(), self.combination_weight(feats))


class SharpeningBlock(_v7._GatedConvFFN):
    """V1:768-787: 1x1 -> depthwise 3x3 -> gelu(a) * b -> 1x1, with a learned two-term skip"""

    def __init__(self, dim_in, dim_out, hidden_features):
        super().__init__(dim_in, hidden_features, dim_out, False)
        self.skip_connect_weight = Parameter(torch.tensor([0.5, 0.5], dtype=torch.float32))

    def forward(self, patchs):
        return self.skip_connect_weight[0] * patchs + self.skip_connect_weight[1] * super().forward(patchs)


def _full_window(n):
    w = np.ones((n, n), dtype=np.int64)
    w[n // 2, n // 2] = 0
    return w


class MultiScaleSequenceDenoiser(nn.Module):
    """V1:790-883: three MixtureGTV blocks in sequence - full 3x3 window (8 edges, G=4, F=6) twice, then the full 5x5 window
    (24 edges, G=4, F=12) - 6 iterations each, each with a weighted skip and followed by a SharpeningBlock."""

    def __init__(self, device):
        super().__init__()
        self.device = device
        z = lambda v: torch.tensor([[v], [0.0], [0.0], [0.0]])
        for tag, win, fts in (("01", 3, 6), ("02", 3, 6), ("03", 5, 12)):
            setattr(self, "skip_connect_weight" + tag, Parameter(torch.tensor([0.1, 0.9], dtype=torch.float32, device=device)))
            setattr(self, "mixtureGLR_block" + tag,
                    MixtureGTV(nchannels_in=3, n_graphs=4, n_node_fts=fts, connection_window=_full_window(win), n_cgd_iters=6,
                               alpha_init=0.5, beta_init=0.1, muy_init=z(0.1), ro_init=z(0.1), gamma_init=z(0.001), device=device))
            setattr(self, "sharp" + tag, SharpeningBlock(3, 3, 24).to(device))

    def forward(self, patchs):
        out = patchs
        for tag in ("01", "02", "03"):
            skip = getattr(self, "skip_connect_weight" + tag)
            out = skip[0] * out + skip[1] * getattr(self, "mixtureGLR_block" + tag)(out)
            out = getattr(self, "sharp" + tag)(out)
        return out
