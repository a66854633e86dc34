"""CPU oracle for the unrolled GLR / GTV restoration blocks.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
file: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs use it, and only as the checker / the timed CPU arm.

What it is
----------
A plain functional restatement, in PyTorch (CPU, any float dtype), of the
arithmetic of the reference's graph-filter blocks.  It is written with three
shift primitives (`shift_clamp`, `shift_reflect`, `shift_zero`) instead of the
reference's pad/stack/slice/conv calls, so it shares no code with the
reference.  PyTorch rather than numpy/C because the path is floating point and
the oracle must also deliver *gradients* (through autograd) for the backward
kernels.

Reference lines restated (V1X0 = exploration/GGTV_GGLR_v1.0/
deep_multiscale_GGLR_GGTV_v1x0.py, byte-identical to LIB/model_GLR_GTV_deep_v13.py):

  window_edges            V1X0:42-53     edge order = row-major (dh,dw) over the mask
  edge_weights            V1X0:146-175   normalise, scale by multiM, dot, softmax
  stats_taps / S / St     V1X0:177-215   depthwise 5-tap conv and its transpose
  op_L                    V1X0:218-228
  op_C / op_Ct            V1X0:452-516   (incl. the "drop" rule of the scatter)
  soft_threshold          V1X0:684-704
  pool2 / unpool2         V1X0:613, 662-665, 676-679
  apply_A                 V1X0:642-682   apply_lightweight_transformer
  mixture_gtvglr_forward  V1X0:707-811
  lowpass_block_forward   V1X0:985-988

Pinning: tests/golden/*.npz were produced by the *reference itself* (imported
from /root/reference by tests/golden/make_golden.py) and tests/test_oracle.py
checks this file against them.  The reference ships no tests or golden vectors
of its own, so those fixtures are the only pin.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch

# ----------------------------------------------------------------------------
# windows
# ----------------------------------------------------------------------------

WINDOWS: Dict[str, List[List[int]]] = {
    "cross3": [[0, 1, 0], [1, 0, 1], [0, 1, 0]],
    "full3": [[1, 1, 1], [1, 0, 1], [1, 1, 1]],
    "small5": [
        [0, 0, 1, 0, 0],
        [0, 1, 1, 1, 0],
        [1, 1, 0, 1, 1],
        [0, 1, 1, 1, 0],
        [0, 0, 1, 0, 0],
    ],
    "full5": [[0 if (i == 2 and j == 2) else 1 for j in range(5)] for i in range(5)],
    "full7": [[0 if (i == 3 and j == 3) else 1 for j in range(7)] for i in range(7)],
}


def window_edges(window) -> List[Tuple[int, int]]:
    """(dh, dw) offsets of the ones of a square 0/1 mask, row-major (V1X0:42-49)."""
    if isinstance(window, str):
        window = WINDOWS[window]
    n = len(window)
    half = n // 2
    out = []
    for i, j in itertools.product(range(n), range(n)):
        if window[i][j] == 1:
            out.append((i - half, j - half))
    return out


# ----------------------------------------------------------------------------
# shift primitives on the last two axes.  result[..., h, w] = x[..., map(h+dh), map(w+dw)]
# ----------------------------------------------------------------------------

def _idx(n: int, d: int, mode: str, device) -> torch.Tensor:
    i = torch.arange(n, device=device) + d
    if mode == "clamp":
        return i.clamp(0, n - 1)
    if mode == "reflect":
        i = torch.where(i < 0, -i, i)
        i = torch.where(i > n - 1, 2 * (n - 1) - i, i)
        return i
    raise ValueError(mode)


def _shift_padded(x: torch.Tensor, dh: int, dw: int, mode: str) -> torch.Tensor:
    # pad by |d| on the side the shift reads from, then take the H x W window that starts at the shift
    H, W = x.shape[-2:]
    if dh == 0 and dw == 0:
        return x
    pads = (max(-dw, 0), max(dw, 0), max(-dh, 0), max(dh, 0))
    lead = x.shape[:-2]
    xp = torch.nn.functional.pad(x.reshape(1, -1, H, W), pads, mode=mode)
    h0, w0 = max(dh, 0), max(dw, 0)
    return xp[..., h0:h0 + H, w0:w0 + W].reshape(*lead, H, W)


def shift_clamp(x: torch.Tensor, dh: int, dw: int) -> torch.Tensor:
    """x[cl(h+dh), cl(w+dw)]"""
    H, W = x.shape[-2:]
    if abs(dh) >= H or abs(dw) >= W:   # window wider than the image: fall back to explicit indices
        return x.index_select(-2, _idx(H, dh, "clamp", x.device)).index_select(-1, _idx(W, dw, "clamp", x.device))
    return _shift_padded(x, dh, dw, "replicate")


def shift_reflect(x: torch.Tensor, dh: int, dw: int) -> torch.Tensor:
    H, W = x.shape[-2:]
    return x.index_select(-2, _idx(H, dh, "reflect", x.device)).index_select(-1, _idx(W, dw, "reflect", x.device))


def shift_zero(x: torch.Tensor, dh: int, dw: int) -> torch.Tensor:
    """result[h,w] = x[h+dh, w+dw] if inside else 0."""
    H, W = x.shape[-2:]
    out = torch.zeros_like(x)
    h0, h1 = max(0, -dh), min(H, H - dh)
    w0, w1 = max(0, -dw), min(W, W - dw)
    if h0 < h1 and w0 < w1:
        out[..., h0:h1, w0:w1] = x[..., h0 + dh:h1 + dh, w0 + dw:w1 + dw]
    return out


# ----------------------------------------------------------------------------
# B.1 edge weights
# ----------------------------------------------------------------------------

def normalize_transform(feat: torch.Tensor, multiM: torch.Tensor) -> torch.Tensor:
    """feat [B,G,F,H,W], multiM [G,F] -> M * feat/max(||feat||_F, 1e-12)  (V1X0:146-157)."""
    nrm = feat.pow(2).sum(dim=2, keepdim=True).sqrt().clamp_min(1e-12)
    return feat / nrm * multiM[None, :, :, None, None]


def edge_weights(feat: torch.Tensor, multiM: torch.Tensor, edges: Sequence[Tuple[int, int]]) -> torch.Tensor:
    """-> w [B,G,E,H,W], softmax over E of <ft[p], ft[cl(p+d_e)]>  (V1X0:160-175)."""
    ft = normalize_transform(feat, multiM)
    sims = [(ft * shift_clamp(ft, dh, dw)).sum(dim=2) for dh, dw in edges]
    return torch.softmax(torch.stack(sims, dim=2), dim=2)


# ----------------------------------------------------------------------------
# B.2 / B.3 stats conv S and "transpose" St
# ----------------------------------------------------------------------------

_TAPS = ((0, 0), (0, 1), (1, 0), (-1, 0), (0, -1))  # c, R, D, U, L


def stats_taps(p1, pa, pb, p3):
    """per-channel tap coefficients (k_c, k_R, k_D, k_U, k_L) from the four parameters."""
    p1, pa, pb, p3 = (t.reshape(-1) for t in (p1, pa, pb, p3))
    return (p1 - pa - pb + 4.0 * p3, pa - p3, pb - p3, -p3, -p3)


def _per_channel(k: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    # x is [B,G,F,H,W] with channel c = g*F+f, or [B,C,H,W]; k is [C] or [1]
    if k.numel() == 1:
        return k.reshape(())
    if x.dim() == 5:
        return k.reshape(1, x.shape[1], x.shape[2], 1, 1)
    return k.reshape(1, -1, 1, 1)


def stats_conv(x, stats, pad_mode: str = "clamp"):
    """S x = sum_t k_t x[map(p+o_t)]   (V1X0:177-195; family A uses reflect, v7:449-467)."""
    ks = stats_taps(*stats)
    sh = shift_clamp if pad_mode == "clamp" else shift_reflect
    out = 0
    for k, (dh, dw) in zip(ks, _TAPS):
        out = out + _per_channel(k, x) * sh(x, dh, dw)
    return out


def stats_conv_transpose(y, stats):
    """St y [q] = sum_t k_t y[q-o_t] [q-o_t inside]  (V1X0:197-215)."""
    ks = stats_taps(*stats)
    out = 0
    for k, (dh, dw) in zip(ks, _TAPS):
        out = out + _per_channel(k, y) * shift_zero(y, -dh, -dw)
    return out


# ----------------------------------------------------------------------------
# B.4 L,  B.5 C,  B.6 Ct
# ----------------------------------------------------------------------------

def op_L(x, w, edges):
    """x [B,G,F,H,W], w [B,G,E,H,W]:  x - sum_e w_e x[cl(p+d_e)]  (V1X0:218-228)."""
    acc = 0
    for e, (dh, dw) in enumerate(edges):
        acc = acc + w[:, :, e:e + 1] * shift_clamp(x, dh, dw)
    return x - acc


def glr_forward(x, w, stats, edges, pad_mode="clamp", use_stats=True):
    """St L S x (V1X0:231-237)."""
    if not use_stats:
        return op_L(x, w, edges)
    return stats_conv_transpose(op_L(stats_conv(x, stats, pad_mode), w, edges), stats)


def op_C_core(s, w, edges):
    """z[:, :, :, e] = w_e (s - s[cl(p+d_e)])   -> [B,G,F,E,H,W]  (V1X0:459-467)."""
    zs = [w[:, :, None, e] * (s - shift_clamp(s, dh, dw)) for e, (dh, dw) in enumerate(edges)]
    return torch.stack(zs, dim=3)


def op_C(x, w, stats, edges, pad_mode="clamp", use_stats=True):
    s = stats_conv(x, stats, pad_mode) if use_stats else x
    return op_C_core(s, w, edges)


def op_Ct_core(z, w, edges):
    """o[q] = sum_e u_e[q] - sum_e u_e[q-d_e][q-d_e inside],  u = w*z  (V1X0:471-513)."""
    u = z * w[:, :, None]
    o = u.sum(dim=3)
    for e, (dh, dw) in enumerate(edges):
        o = o - shift_zero(u[:, :, :, e], -dh, -dw)
    return o


def op_Ct(z, w, stats, edges, use_stats=True):
    o = op_Ct_core(z, w, edges)
    return stats_conv_transpose(o, stats) if use_stats else o


def gtv_forward(x, w, stats, edges, pad_mode="clamp", use_stats=True):
    """Ct C x (V1X0:518-523)."""
    return op_Ct(op_C(x, w, stats, edges, pad_mode, use_stats), w, stats, edges, use_stats)


# ----------------------------------------------------------------------------
# B.7 soft threshold, B.8 pooling
# ----------------------------------------------------------------------------

def soft_threshold(t, thr):
    """t [B,G,F,E,H,W], thr [G] (already exp'd).  strict inequalities (V1X0:684-704)."""
    g = thr[None, :, None, None, None, None]
    zero = torch.zeros((), dtype=t.dtype, device=t.device)
    return torch.where(t < -g, t + g, zero) + torch.where(t > g, t - g, zero)


def pool2(x):
    """mean over aligned 2x2 blocks (V1X0:613, 662-665)."""
    return 0.25 * (x[..., 0::2, 0::2] + x[..., 0::2, 1::2] + x[..., 1::2, 0::2] + x[..., 1::2, 1::2])


def unpool2(x):
    """conv_transpose2d with the 0.25 kernel: 0.25*x copied to each 2x2 block (V1X0:676-679)."""
    return 0.25 * x.repeat_interleave(2, dim=-2).repeat_interleave(2, dim=-1)


# ----------------------------------------------------------------------------
# V1X0 block
# ----------------------------------------------------------------------------

@dataclass
class OpParams:
    """parameters of one GLRFast / GTVFast module."""
    p1: torch.Tensor
    pa: torch.Tensor
    pb: torch.Tensor
    p3: torch.Tensor
    multiM: torch.Tensor

    @property
    def stats(self):
        return (self.p1, self.pa, self.pb, self.p3)


def op_params_from_state(sd: Dict[str, torch.Tensor], prefix: str) -> OpParams:
    return OpParams(
        sd[prefix + "stats_kernel_p01"], sd[prefix + "stats_kernel_p02a"],
        sd[prefix + "stats_kernel_p02b"], sd[prefix + "stats_kernel_p03"],
        sd[prefix + "multiM"],
    )


def _conv1x1(x, w):
    # w [Co,Ci,1,1]
    return torch.einsum("oi,bihw->bohw", w[:, :, 0, 0], x)


def _conv2x2s2(x, w):
    # w [Co,Ci,2,2], stride 2, no padding (cross-correlation)
    out = 0
    for i in range(2):
        for j in range(2):
            out = out + torch.einsum("oi,bihw->bohw", w[:, :, i, j], x[:, :, i::2, j::2])
    return out


def mixture_gtvglr_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "",
                           return_intermediates: bool = False):
    """MixtureGTVGLR.forward (V1X0:707-811) from a state-dict of that module.

    x [B,C,H,W]; G, F inferred from alphaCGD / multiM shapes.
    """
    g = lambda k: sd[prefix + k]
    edges = window_edges("cross3")
    T0 = op_params_from_state(sd, prefix + "GTVmodule00.")
    L0 = op_params_from_state(sd, prefix + "GLRmodule00.")
    T1 = op_params_from_state(sd, prefix + "GTVmodule01.")
    L1 = op_params_from_state(sd, prefix + "GLRmodule01.")
    G, F = T0.multiM.shape
    B, C, H, W = x.shape
    assert C == G * F and H % 2 == 0 and W % 2 == 0

    alpha, beta = g("alphaCGD"), g("betaCGD")
    mu0, ro0, ga0 = g("muys00").exp(), g("ro00").exp(), g("gamma00").exp()
    mu1, ro1, ga1 = g("muys01").exp(), g("ro01").exp(), g("gamma01").exp()
    bc = lambda v: v[None, :, None, None, None]

    # features -> edge weights, fine and coarse (V1X0:712-733)
    f0 = _conv1x1(x, g("patchs_features_extraction00.0.weight"))
    f1 = _conv1x1(_conv2x2s2(x, g("patchs_features_extraction01.0.weight")),
                  g("patchs_features_extraction01.1.weight"))
    v5 = lambda t: t.reshape(B, G, F, t.shape[-2], t.shape[-1])
    wT0 = edge_weights(v5(f0[:, :C]), T0.multiM, edges)
    wL0 = edge_weights(v5(f0[:, C:]), L0.multiM, edges)
    wT1 = edge_weights(v5(f1[:, :C]), T1.multiM, edges)
    wL1 = edge_weights(v5(f1[:, C:]), L1.multiM, edges)

    def A(z):  # V1X0:642-682
        out = z + bc(mu0) * glr_forward(z, wL0, L0.stats, edges) + bc(ro0) * gtv_forward(z, wT0, T0.stats, edges)
        zc = pool2(z)
        tc = bc(mu1) * glr_forward(zc, wL1, L1.stats, edges) + bc(ro1) * gtv_forward(zc, wT1, T1.stats, edges)
        return out + unpool2(tc)

    y = v5(x)
    # pass A (V1X0:738-753)
    bA = (y + bc(ro0) * op_Ct(op_C(y, wT0, T0.stats, edges), wT0, T0.stats, edges)
          + bc(ro1) * unpool2(op_Ct(op_C(pool2(y), wT1, T1.stats, edges), wT1, T1.stats, edges)))
    x1 = bA + bc(alpha[0]) * (bA - A(bA))
    # threshold step (V1X0:757-781)
    t0 = op_C(x1, wT0, T0.stats, edges)
    t1 = op_C(pool2(x1), wT1, T1.stats, edges)
    e0, e1 = soft_threshold(t0, ga0), soft_threshold(t1, ga1)
    bB = (y + bc(ro0) * op_Ct(e0 - (t0 - e0), wT0, T0.stats, edges)
          + bc(ro1) * unpool2(op_Ct(e1 - (t1 - e1), wT1, T1.stats, edges)))
    # two more iterations (V1X0:784-790)
    r1 = bB - A(x1)
    x2 = x1 + bc(alpha[1]) * r1
    r2 = bB - A(x2)
    u2 = r2 + bc(beta[2]) * r1
    x3 = x2 + bc(alpha[2]) * u2
    out = x3.reshape(B, C, H, W)
    if return_intermediates:
        return out, dict(wT0=wT0, wL0=wL0, wT1=wT1, wL1=wL1, bA=bA, x1=x1, bB=bB, r1=r1, x2=x2)
    return out


def lowpass_block_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "") -> torch.Tensor:
    """LocalLowpassFilteringBlock.forward (V1X0:985-988)."""
    s = sd[prefix + "skip_weight"]
    return s[0] * x + s[1] * mixture_gtvglr_forward(sd, x, prefix + "local_filter.")


def lowpass_block_fwd_bwd(sd: Dict[str, torch.Tensor], x: torch.Tensor, gout: torch.Tensor, prefix: str = ""):
    """forward + autograd backward; returns (out, gx, {param: grad})."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if k.startswith(prefix)}
    xx = x.detach().clone().requires_grad_(True)
    out = lowpass_block_forward(leaf, xx, prefix)
    names = list(leaf.keys())
    grads = torch.autograd.grad(out, [xx] + [leaf[k] for k in names], gout, allow_unused=True)
    pg = {k: (gr if gr is not None else torch.zeros_like(leaf[k])) for k, gr in zip(names, grads[1:])}
    return out.detach(), grads[0], pg


# ----------------------------------------------------------------------------
# parameter randomisation used by tests / fixtures
# ----------------------------------------------------------------------------

def randomize_block_state(sd: Dict[str, torch.Tensor], seed: int, prefix: str = "") -> Dict[str, torch.Tensor]:
    """Move every parameter of a LocalLowpassFilteringBlock state-dict away from its init so that
    all graph terms contribute (default init makes the block ~identity)."""
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if not k.startswith(prefix):
            out[k] = v
            continue
        name = k[len(prefix):]
        r = lambda *s: torch.rand(*s, generator=gen, dtype=torch.float64)
        n = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float64)
        if name.endswith(("muys00", "muys01", "ro00", "ro01")):
            nv = torch.log(0.01 + 0.04 * r(*v.shape))
        elif name.endswith(("gamma00", "gamma01")):
            nv = torch.log(0.01 + 0.49 * r(*v.shape))
        elif name.endswith("multiM"):
            nv = 1.0 + 0.5 * n(*v.shape)
        elif "stats_kernel_p" in name:
            nv = v.double() + 0.1 * n(*v.shape)
        elif name.endswith(("alphaCGD", "betaCGD")):
            nv = v.double() + 0.05 * n(*v.shape)
        elif name.endswith("skip_weight"):
            nv = v.double() + 0.1 * n(*v.shape)
        else:  # projection weights: keep their init scale, new values
            nv = v.double() + 0.05 * n(*v.shape)
        out[k] = nv.to(v.dtype)
    return out


# ----------------------------------------------------------------------------
# older family ("family A"): MixtureGTV of LIB/model_GLR_GTV_deep_v7.py:936-1016
# ----------------------------------------------------------------------------

def unrolled_admm_solve(sd: Dict[str, torch.Tensor], y: torch.Tensor, wT: torch.Tensor, wL: torch.Tensor, edges,
                        schedule=(2, 2), prefix: str = "") -> torch.Tensor:
    """Loop-generalised restatement of the older family's solver (v7:964-1000 is schedule (2,2), v1:627-668 is (2,4),
    v0:644-682 is (2,2,2)): y [B,G,c,H,W]; each entry of `schedule` is one inner solve of that many momentum iterations
    restarted from its right-hand side; between solves one soft-threshold with the carried dual `bias`."""
    g = lambda k: sd[prefix + k]
    T = op_params_from_state(sd, prefix + "GTVmodule00.")
    Lp = op_params_from_state(sd, prefix + "GLRmodule00.")
    bc = lambda v: v[None, :, None, None, None]
    ro, mu, gam = g("ro00"), g("muys00"), g("gamma00").exp()
    alpha, beta = g("alphaCGD"), g("betaCGD")
    C = lambda z: op_C(z, wT, T.stats, edges, "reflect")
    Ct = lambda e: op_Ct(e, wT, T.stats, edges)
    A = lambda z: (z + bc(mu) * glr_forward(z, wL, Lp.stats, edges, "reflect") + bc(ro) * Ct(C(z)))
    eps, bias, k, out = C(y), torch.zeros(()), 0, None
    for n_solve, n_it in enumerate(schedule):
        if n_solve > 0:
            t = C(out)
            eps = soft_threshold(t + bias, gam)
            bias = bias + (t - eps)
        rhs = Ct(eps - bias) * bc(ro) + y
        out, upd = rhs, None
        for _ in range(n_it):
            r = rhs - A(out)
            upd = r if upd is None else r + bc(beta[k]) * upd
            out = out + bc(alpha[k]) * upd
            k += 1
    return out


def mixture_gtv_solver(sd: Dict[str, torch.Tensor], patchs: torch.Tensor, feats: torch.Tensor, dc_term: torch.Tensor,
                       score: torch.Tensor, window="small5", schedule=(2, 2), prefix: str = "") -> torch.Tensor:
    """The graph part of MixtureGTV.forward given the CNN outputs: feats [B, G*F+12, H, W] (features), dc_term
    [B,3,H,W], score [B,G,H,W] (soft-maxed mixture weights).  Scalar stats parameters, reflect-padded S, the RGB
    signal broadcast over the G graphs, raw ro/muys, log gamma (v7:936-1016)."""
    edges = window_edges(window)
    T = op_params_from_state(sd, prefix + "GTVmodule00.")
    Lp = op_params_from_state(sd, prefix + "GLRmodule00.")
    G, F = T.multiM.shape
    B, _, H, W = patchs.shape
    gfeat = feats[:, :-12].reshape(B, G, F, H, W)
    wT, wL = edge_weights(gfeat, T.multiM, edges), edge_weights(gfeat, Lp.multiM, edges)
    y = (patchs - dc_term)[:, None].expand(B, G, 3, H, W)
    out = unrolled_admm_solve(sd, y, wT, wL, edges, schedule, prefix)
    return (out * score[:, :, None]).sum(dim=1) + dc_term
