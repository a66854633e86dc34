"""Drop-in for the reference's older "multiblock" family, LIB/model_GLR_GTV_deep_v7.py (V7), the model that
scripts/run_lightformer_GGTV_GGLR_multiblocks.py:32,156 trains.

Same class names, constructor arguments and `state_dict` keys as V7.  The graph operators run the hand-written
sm_100a kernels (any connection window, reflect-padded S with scalar parameters, RGB signal broadcast over the
G graphs, split-Bregman pass with a carried dual `bias`, mixture weighting); the feature-extraction CNN and the DC
estimator around them are out of the hot path and are plain PyTorch layers with the reference's parameter layout.

  GLRFast / GTVFast (family A)   V7:274-509, 514-781
  MixtureGTV                     V7:802-1016
  MultiScaleSequenceDenoiser     V7:1019-1087
"""
import itertools

import numpy as np
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import ops
from .ops_mixture import mixture

_REFLECT = 1


# ----------------------------------------------------------------------------------------------------
# out-of-scope CNN pieces (parameter layout of V7:13-125, 195-269, 784-799)
# ----------------------------------------------------------------------------------------------------
class CustomLayerNorm(nn.Module):
    def __init__(self, nchannels):
        super().__init__()
        self.nchannels = nchannels
        self.weighted_transform = nn.Conv2d(nchannels, nchannels, kernel_size=1, groups=nchannels, bias=False)

    def forward(self, x):
        return self.weighted_transform(x / torch.sqrt(x.var(dim=1, keepdim=True, correction=1) + 1e-5))


class _GatedConvFFN(nn.Module):
    """1x1 -> depthwise 3x3 -> gelu(a)*b -> 1x1 (FeedForward V7:29-47 and DCestimator V7:784-799 share it)."""

    def __init__(self, dim_in, hidden, dim_out, bias):
        super().__init__()
        self.project_in = nn.Conv2d(dim_in, 2 * hidden, kernel_size=1, bias=bias)
        self.dwconv = nn.Conv2d(2 * hidden, 2 * hidden, kernel_size=3, padding=1, groups=2 * hidden, bias=bias)
        self.project_out = nn.Conv2d(hidden, dim_out, kernel_size=1, bias=bias)

    def forward(self, x):
        a, b = self.dwconv(self.project_in(x)).chunk(2, dim=1)
        return self.project_out(nn.functional.gelu(a) * b)


class FeedForward(_GatedConvFFN):
    def __init__(self, dim, ffn_expansion_factor, bias):
        super().__init__(dim, int(dim * ffn_expansion_factor), dim, bias)


class DCestimator(_GatedConvFFN):
    def __init__(self, dim_in, dim_out, hidden_features):
        super().__init__(dim_in, hidden_features, dim_out, False)


class FFBlock(nn.Module):
    def __init__(self, dim, ffn_expansion_factor, bias):
        super().__init__()
        self.norm = CustomLayerNorm(dim)
        self.skip_connect_weight_final = Parameter(torch.tensor([0.5, 0.5], dtype=torch.float32))
        self.ffn = FeedForward(dim, ffn_expansion_factor, bias)

    def forward(self, x):
        return self.skip_connect_weight_final[0] * x + self.skip_connect_weight_final[1] * self.ffn(self.norm(x))


class OverlapPatchEmbed(nn.Module):
    def __init__(self, in_c=3, embed_dim=48, bias=False):
        super().__init__()
        self.proj = nn.Conv2d(in_c, embed_dim, kernel_size=3, padding=1, bias=bias)

    def forward(self, x):
        return self.proj(x)


class Downsample(nn.Module):
    def __init__(self, n_feat):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(n_feat, n_feat // 2, kernel_size=3, padding=1, bias=False), nn.PixelUnshuffle(2))

    def forward(self, x):
        return self.body(x)


class Upsample(nn.Module):
    def __init__(self, n_feat):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(n_feat, n_feat * 2, kernel_size=3, padding=1, bias=False), nn.PixelShuffle(2))

    def forward(self, x):
        return self.body(x)


class FeatureExtraction(nn.Module):
    """two-level conv U-Net (V7:195-269); returns a one-element list like the reference."""

    def __init__(self, inp_channels=3, out_channels=48, dim=48, num_blocks=[1, 2, 2, 4], num_refinement_blocks=4,
                 ffn_expansion_factor=2.66, bias=False):
        super().__init__()
        blocks = lambda d, n: nn.Sequential(*[FFBlock(d, ffn_expansion_factor, bias) for _ in range(n)])
        self.patch_embed = OverlapPatchEmbed(inp_channels, dim)
        self.encoder_level1 = blocks(dim, num_blocks[0])
        self.down1_2 = Downsample(dim)
        self.encoder_level2 = blocks(2 * dim, num_blocks[1])
        self.up2_1 = Upsample(2 * dim)
        self.decoder_level1 = blocks(2 * dim, num_blocks[0])
        self.refinement = blocks(2 * dim, num_refinement_blocks)
        self.output = nn.Conv2d(2 * dim, out_channels, kernel_size=3, padding=1, bias=bias)

    def forward(self, inp_img):
        e1 = self.encoder_level1(self.patch_embed(inp_img))
        lat = self.encoder_level2(self.down1_2(e1))
        d1 = self.decoder_level1(torch.cat([self.up2_1(lat), e1], 1))
        return [self.output(self.refinement(d1))]


# ----------------------------------------------------------------------------------------------------
# graph operators, family A
# ----------------------------------------------------------------------------------------------------
class _GraphOperatorA(nn.Module):
    """V7:274-370 / 514-610: window from a 0/1 mask, SCALAR stats_kernel_p* shared by all channels, multiM [G,F]."""

    def __init__(self, n_channels, n_node_fts, n_graphs, connection_window, device, M_diag_init=0.4):
        super().__init__()
        self.device = device
        self.n_channels = n_channels
        self.n_node_fts = n_node_fts
        self.n_graphs = n_graphs
        mask = np.asarray(connection_window)
        self.n_edges = int((mask == 1).sum())
        self.connection_window = mask
        self.buffer_size = int(mask.sum())
        half = mask.shape[0] // 2
        offs = np.arange(mask.shape[0]) - half
        self.edge_delta = np.array([d for d, on in zip(itertools.product(offs, offs), mask.reshape(-1)) if on == 1], dtype=np.int32)
        self.pad_dim_hw = np.abs(self.edge_delta.min(axis=0))
        self._edges_flat = ops.flat_edges(self.edge_delta.tolist())
        for name, init in (("stats_kernel_p01", 1.0), ("stats_kernel_p02a", 0.5), ("stats_kernel_p02b", 0.5), ("stats_kernel_p03", 0.5)):
            setattr(self, name, Parameter(torch.full((1,), init, dtype=torch.float32, device=device)))
        self.multiM = Parameter(torch.full((n_graphs, n_node_fts), float(M_diag_init), dtype=torch.float32, device=device))

    def _stats(self):
        return (self.stats_kernel_p01, self.stats_kernel_p02a, self.stats_kernel_p02b, self.stats_kernel_p03)

    def get_neighbors_pixels(self, img_features):
        b, c, h, w = img_features.shape
        return ops.gather_neighbors(img_features.reshape(b, 1, c, h, w), self._edges_flat).reshape(b, c, self.n_edges, h, w)

    def normalize_and_transform_features(self, img_features):
        b, g, f, h, w = img_features.shape
        return ops.normalize_transform(img_features, self.multiM).reshape(b, g * f, h, w)

    def extract_edge_weights(self, img_features):
        w = ops.edge_weights(img_features, self.multiM, self._edges_flat)
        return w, w.sum(dim=2)

    def stats_conv(self, patchs):          # reflect padding, V7:449-467
        return ops.stats_conv(patchs, *self._stats(), _REFLECT)

    def stats_conv_transpose(self, patchs):
        return ops.stats_conv_t(patchs, *self._stats(), _REFLECT)

    def _over_graphs(self, signal, weights):
        """the reference lets a [B,1,c,H,W] signal broadcast against G graphs (V7:966); the kernels want it explicit"""
        if signal.shape[1] == 1 and weights.shape[1] != 1:
            signal = signal.expand(-1, weights.shape[1], -1, -1, -1)
        return signal.contiguous()


class GLRFast(_GraphOperatorA):
    def op_L_norm(self, img_signals, edge_weights, node_degree):
        return ops.op_L(self._over_graphs(img_signals, edge_weights), edge_weights, self._edges_flat)

    def forward(self, patchs, edge_weights, node_degree):
        return self.stats_conv_transpose(self.op_L_norm(self.stats_conv(patchs), edge_weights, node_degree))


class GTVFast(_GraphOperatorA):
    def op_C(self, img_signals, edge_weights, node_degree):
        s = self.stats_conv(img_signals.contiguous())
        return ops.op_C(self._over_graphs(s, edge_weights), edge_weights, self._edges_flat)

    def op_C_transpose(self, edge_signals, edge_weights, node_degree):
        return self.stats_conv_transpose(ops.op_Ct(edge_signals, edge_weights, self._edges_flat))

    def forward(self, patchs, edge_weights, node_degree):
        return self.op_C_transpose(self.op_C(patchs, edge_weights, node_degree), edge_weights, node_degree)


class MixtureGTV(nn.Module):
    """V7:802-1016: one split-Bregman pass with a carried dual, four unrolled momentum iterations, mixture output."""

    def __init__(self, nchannels_in, n_graphs, n_node_fts, n_cnn_fts, connection_window, n_cgd_iters, alpha_init, beta_init,
                 muy_init, ro_init, gamma_init, device):
        super().__init__()
        self.device = device
        self.n_graphs, self.n_node_fts = n_graphs, n_node_fts
        self.n_total_fts = n_graphs * n_node_fts
        self.n_cnn_fts, self.n_levels, self.n_cgd_iters = n_cnn_fts, 4, n_cgd_iters
        self.nchannels_in, self.connection_window = nchannels_in, connection_window
        vec = lambda v: Parameter(torch.ones(n_graphs, dtype=torch.float32, device=device) * v)
        self.alphaCGD = Parameter(torch.full((n_cgd_iters, n_graphs), float(alpha_init), dtype=torch.float32, device=device))
        self.betaCGD = Parameter(torch.full((n_cgd_iters, n_graphs), float(beta_init), dtype=torch.float32, device=device))
        self.patchs_features_extraction = FeatureExtraction(
            inp_channels=3, out_channels=self.n_total_fts + 12, dim=n_cnn_fts, num_blocks=[4, 3, 3],
            num_refinement_blocks=4, ffn_expansion_factor=2.6666, bias=False).to(device)
        self.combination_weight = nn.Sequential(nn.Conv2d(self.n_total_fts, n_graphs, kernel_size=1, bias=False),
                                                nn.Softmax(dim=1)).to(device)
        self.dc_estimator = DCestimator(12, 3, 12 * 2).to(device)
        scalar = lambda t: torch.as_tensor(t, dtype=torch.float32).reshape(-1)[0].to(device)
        self.ro00 = vec(scalar(ro_init[0]))                           # raw, not log (V7:862-865)
        self.gamma00 = vec(torch.log(scalar(gamma_init[0])))          # log domain (V7:866-869)
        self.GTVmodule00 = GTVFast(nchannels_in, n_node_fts, n_graphs, connection_window, device, M_diag_init=1.0)
        self.muys00 = vec(scalar(muy_init[0]))
        self.GLRmodule00 = GLRFast(nchannels_in, n_node_fts, n_graphs, connection_window, device, M_diag_init=1.0)

    def apply_lightweight_transformer(self, patchs, list_graph_weightGTV, list_graph_weightGLR):
        bc = lambda v: v[None, :, None, None, None]
        z = patchs.contiguous()
        return (z + bc(self.muys00) * self.GLRmodule00(z, *list_graph_weightGLR[0])
                + bc(self.ro00) * self.GTVmodule00(z, *list_graph_weightGTV[0]))

    def soft_threshold(self, delta, gamma):
        return ops.soft_threshold(delta, gamma)

    def unrolled_solve(self, y, wT, wL, schedule=(2, 2)):
        """The unrolled split-Bregman solver on a signal y [B,1 or G,c,H,W] for given edge weights.
        `schedule` lists the momentum iterations of each inner solve; between two solves comes one soft-threshold with
        the dual update.  (2, 2) is V7:964-1000, (2, 4) is v1:627-668, (2, 2, 2) is v0:644-682; alphaCGD / betaCGD need
        sum(schedule) rows.  Each solve restarts from its right-hand side, its first iteration has no momentum term."""
        T, bc = self.GTVmodule00, (lambda v: v[None, :, None, None, None])
        a, be = self.alphaCGD[:, None, :, None, None, None], self.betaCGD[:, None, :, None, None, None]
        A = lambda z: self.apply_lightweight_transformer(z, [wT], [wL])
        if sum(schedule) > self.alphaCGD.shape[0]:
            raise ValueError(f"schedule {schedule} needs {sum(schedule)} rows of alphaCGD, have {self.alphaCGD.shape[0]}")
        y = T._over_graphs(y, wT[0])
        eps, bias, k, out = T.op_C(y, *wT), None, 0, None
        for n_solve, n_it in enumerate(schedule):
            if n_solve > 0:                                   # threshold + dual update (V7:982-986)
                t = T.op_C(out, *wT)
                eps = self.soft_threshold(t if bias is None else t + bias, torch.exp(self.gamma00))
                bias = (t - eps) if bias is None else bias + (t - eps)
            rhs = T.op_C_transpose(eps if bias is None else eps - bias, *wT) * bc(self.ro00) + y
            out, upd = rhs, None
            for _ in range(n_it):
                r = rhs - A(out)
                upd = r if upd is None else r + be[k] * upd
                out = out + a[k] * upd
                k += 1
        return out

    def forward(self, patchs):
        feats = self.patchs_features_extraction(patchs)[0]
        b, _, h, w = feats.shape
        gfeat = feats[:, :-12].reshape(b, self.n_graphs, self.n_node_fts, h, w)
        wT, wL = self.GTVmodule00.extract_edge_weights(gfeat), self.GLRmodule00.extract_edge_weights(gfeat)
        dc_term = self.dc_estimator(feats[:, -12:])
        out = self.unrolled_solve((patchs - dc_term)[:, None], wT, wL, schedule=(2, 2))
        score = self.combination_weight(feats[:, :-12])
        return mixture(out.contiguous(), score) + dc_term


class MultiScaleSequenceDenoiser(nn.Module):
    """V7:1019-1087: one MixtureGTV block (G=24, F=3, 5x5-small window, 4 iterations) with a weighted skip."""

    def __init__(self, device):
        super().__init__()
        self.device = device
        window = np.array([[0, 0, 1, 0, 0], [0, 1, 1, 1, 0], [1, 1, 0, 1, 1], [0, 1, 1, 1, 0], [0, 0, 1, 0, 0]])
        self.skip_connect_weight03 = Parameter(torch.tensor([0.1, 0.9], dtype=torch.float32, device=device))
        z = lambda v: torch.tensor([[v], [0.0], [0.0], [0.0]])
        self.mixtureGLR_block03 = MixtureGTV(nchannels_in=3, n_graphs=24, n_node_fts=3, n_cnn_fts=128, connection_window=window,
                                             n_cgd_iters=4, alpha_init=0.5, beta_init=0.1, muy_init=z(0.1), ro_init=z(0.1),
                                             gamma_init=z(0.001), device=device)

    def forward(self, patchs):
        return self.skip_connect_weight03[0] * patchs + self.skip_connect_weight03[1] * self.mixtureGLR_block03(patchs)
