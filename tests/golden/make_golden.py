"""Generate golden vectors by running the REFERENCE ITSELF (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes small .npz fixtures next to this file.  Each fixture stores the inputs, the full
state-dict used, and the outputs / gradients the reference produced, in float64 (reference
run in float64 so that the fixture is the exact-arithmetic answer to ~1e-15) and the
float32 result of the reference for the noise-floor check.

Reference entry points exercised (V1X0 = exploration/GGTV_GGLR_v1.0/deep_multiscale_GGLR_GGTV_v1x0.py):
  LocalLowpassFilteringBlock.forward  V1X0:985-988 (-> MixtureGTVGLR.forward V1X0:707-811)
  GLRFast/GTVFast.extract_edge_weights, stats_conv, stats_conv_transpose, op_L_norm,
  op_C, op_C_transpose, forward       V1X0:146-237, 377-523
  MixtureGTVGLR.soft_threshold        V1X0:684-704
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/exploration/GGTV_GGLR_v1.0")

import deep_multiscale_GGLR_GGTV_v1x0 as ref  # noqa: E402  (the reference, unmodified)
from oracle.glr_gtv_oracle import randomize_block_state  # noqa: E402  (only the param randomiser)

_CONST_NAMES = ("stats_kernel01", "stats_kernel02a", "stats_kernel02b", "stats_kernel03")


def to_double(block):
    """module.double() does not move the plain-tensor constants: cast them by hand."""
    block.double()
    for m in block.modules():
        for n in _CONST_NAMES:
            if hasattr(m, n):
                setattr(m, n, getattr(m, n).double())
        if hasattr(m, "scaling_kernel01"):
            m.scaling_kernel01 = m.scaling_kernel01.double()
    return block


def block_case(name, dim, ngraphs, B, H, W, seed):
    torch.manual_seed(seed)
    blk = ref.LocalLowpassFilteringBlock(dim=dim, nsubnets=1, ngraphs=ngraphs)
    sd = randomize_block_state({k: v.detach().clone() for k, v in blk.state_dict().items()}, seed + 1)
    gen = torch.Generator().manual_seed(seed + 2)
    x = torch.randn(B, dim, H, W, generator=gen, dtype=torch.float64)
    gout = torch.randn(B, dim, H, W, generator=gen, dtype=torch.float64)

    # float32 reference run (noise floor)
    blk.load_state_dict(sd)
    with torch.no_grad():
        out32 = blk(x.float()).double()

    # float64 reference run with autograd
    blk = to_double(blk)
    blk.load_state_dict({k: v.double() for k, v in sd.items()})
    xx = x.clone().requires_grad_(True)
    out = blk(xx)
    names = [k for k, _ in blk.named_parameters()]
    params = [p for _, p in blk.named_parameters()]
    grads = torch.autograd.grad(out, [xx] + params, gout, allow_unused=True)
    rec = {"x": x.numpy(), "gout": gout.numpy(), "out": out.detach().numpy(), "out32": out32.numpy(),
           "gx": grads[0].numpy(), "meta": np.array([dim, ngraphs, B, H, W])}
    for k, v in sd.items():
        rec["sd." + k] = v.double().numpy()
    for k, gr in zip(names, grads[1:]):
        rec["grad." + k] = (gr if gr is not None else torch.zeros_like(dict(zip(names, params))[k])).numpy()
    rel = float((out.detach() - out32).norm() / out.detach().norm())
    chg = float((out.detach() - x).norm() / x.norm())
    print(f"{name}: |out-x|/|x|={chg:.3f}  ref fp32-vs-fp64 rel={rel:.2e}")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)


def operator_case(name, F_, G, B, H, W, seed):
    """per-operator goldens from the public methods of GLRFast / GTVFast."""
    torch.manual_seed(seed)
    glr = ref.GLRFast(n_node_fts=F_, n_graphs=G, M_diag_init=1.0)
    gtv = ref.GTVFast(n_node_fts=F_, n_graphs=G, M_diag_init=1.0)
    mix = ref.MixtureGTVGLR.__new__(ref.MixtureGTVGLR)  # only for soft_threshold (stateless method)
    gen = torch.Generator().manual_seed(seed)
    C = F_ * G
    rec = {"meta": np.array([F_, G, B, H, W])}
    for mod, tag in ((glr, "glr"), (gtv, "gtv")):
        to_double(mod)
        with torch.no_grad():
            for pn in ("stats_kernel_p01", "stats_kernel_p02a", "stats_kernel_p02b", "stats_kernel_p03"):
                p = getattr(mod, pn)
                p.add_(0.2 * torch.randn(p.shape, generator=gen, dtype=torch.float64))
                rec[f"{tag}.{pn}"] = p.detach().numpy().copy()
            mod.multiM.copy_(1.0 + 0.5 * torch.randn(mod.multiM.shape, generator=gen, dtype=torch.float64))
            rec[f"{tag}.multiM"] = mod.multiM.detach().numpy().copy()
    feat = torch.randn(B, G, F_, H, W, generator=gen, dtype=torch.float64)
    xs = torch.randn(B, G, F_, H, W, generator=gen, dtype=torch.float64)
    with torch.no_grad():
        w_glr, deg = glr.extract_edge_weights(feat)
        w_gtv, _ = gtv.extract_edge_weights(feat * 0.7 + 0.1)
        zs = torch.randn(B, G, F_, 4, H, W, generator=gen, dtype=torch.float64)
        thr = 0.3 + torch.rand(G, generator=gen, dtype=torch.float64)
        rec.update({
            "feat": feat.numpy(), "x": xs.numpy(), "z": zs.numpy(), "thr": thr.numpy(),
            "w_glr": w_glr.numpy(), "w_gtv": w_gtv.numpy(),
            "glr.S": glr.stats_conv(xs).numpy(), "glr.St": glr.stats_conv_transpose(xs).numpy(),
            "glr.L": glr.op_L_norm(xs, w_glr, deg).numpy(), "glr.fwd": glr(xs, w_glr, deg).numpy(),
            "gtv.S": gtv.stats_conv(xs).numpy(), "gtv.St": gtv.stats_conv_transpose(xs).numpy(),
            "gtv.C": gtv.op_C(xs, w_gtv, deg).numpy(),
            "gtv.Ct": gtv.op_C_transpose(zs.clone(), w_gtv, deg).numpy(),
            "gtv.fwd": gtv(xs, w_gtv, deg).numpy(),
            "soft": ref.MixtureGTVGLR.soft_threshold(mix, zs, thr).numpy(),
        })
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(f"{name}: written")


if __name__ == "__main__":
    torch.set_num_threads(8)
    # F=6 (scales 0/1 of the shipped config) and F=12 (scales 2/3); odd-ish sizes, non-square
    block_case("block_f6_g2", dim=12, ngraphs=2, B=2, H=12, W=20, seed=11)
    block_case("block_f12_g2", dim=24, ngraphs=2, B=1, H=10, W=14, seed=23)
    block_case("block_f6_g4_tiny", dim=24, ngraphs=4, B=1, H=2, W=4, seed=37)   # smallest legal: coarse grid 1x2
    operator_case("ops_f6_g2", F_=6, G=2, B=2, H=7, W=9, seed=5)
    operator_case("ops_f3_g3_small", F_=3, G=3, B=1, H=2, W=3, seed=7)
