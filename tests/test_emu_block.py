"""Fused block kernels (csrc/block_fwd.cu, block_bwd.cu) in the g++ emulation build against the oracle."""
import pytest
import torch

from oracle import glr_gtv_oracle as O
from imagerestoration_development_unrolling_b200 import _lib as L
from tests import emu_harness as E
from tests.util import (rel, random_block_state, block_structs, alloc_saved, oracle_features, alloc_grads,
                        grads_to_state_names)

# (dim, ngraphs, B, H, W): partial tiles, multi-tile (tile = 32x32), smallest legal size
# widths that are multiples of 8 take the branch-free kernels, the others the generic ones
CASES = [(12, 2, 2, 12, 20), (24, 2, 1, 34, 66), (24, 4, 1, 2, 4), (6, 1, 1, 64, 32), (12, 2, 1, 40, 36),
         (12, 2, 1, 40, 24), (24, 2, 1, 34, 64), (12, 2, 2, 16, 8), (6, 1, 1, 70, 72)]


@pytest.mark.parametrize("case", CASES)
def test_block_forward(case):
    dim, G, B, H, W = case
    F = dim // G
    sd = random_block_state(dim, G, seed=100 + H)
    x = torch.randn(B, dim, H, W, generator=torch.Generator().manual_seed(H * W))
    ref_out, inter = O.mixture_gtvglr_forward({k: v.double() for k, v in sd.items()}, x.double(), "local_filter.",
                                              return_intermediates=True)
    s = sd["skip_weight"].double()
    ref_out = s[0] * x.double() + s[1] * ref_out
    f0, f1 = oracle_features(sd, x)
    p, keep = block_structs(sd)
    sv, saved = alloc_saved(B, G, F, H, W)
    out = torch.empty_like(x)
    E.call("glrgtv_block_fwd", L.make_shape(B, G, F, H, W), p, x, f0, f1, out, sv, None)
    for n in ("wT0", "wL0", "wT1", "wL1"):
        assert rel(saved[n], inter[n]) < 5e-6, n
    for n in ("bA", "x1", "bB", "r1", "x2"):
        assert rel(saved[n], inter[n]) < 2e-5, (n, rel(saved[n], inter[n]))
    assert rel(out, ref_out) < 1e-5


@pytest.mark.parametrize("case", CASES)
def test_block_backward(case):
    dim, G, B, H, W = case
    F = dim // G
    sd = random_block_state(dim, G, seed=200 + H)
    gen = torch.Generator().manual_seed(H * W + 1)
    x = torch.randn(B, dim, H, W, generator=gen)
    gout = torch.randn(B, dim, H, W, generator=gen)
    sd64 = {k: v.double() for k, v in sd.items()}
    _, gx_ref, pg_ref = O.lowpass_block_fwd_bwd(sd64, x.double(), gout.double())

    # projections with autograd (the product does these with library GEMMs through torch as well)
    xx = x.clone().requires_grad_(True)
    pw = {k: sd[k].clone().requires_grad_(True) for k in sd if "patchs_features_extraction" in k}
    f0, f1 = oracle_features({**sd, **pw}, xx)
    p, keep = block_structs(sd)
    sv, saved = alloc_saved(B, G, F, H, W)
    out = torch.empty_like(x)
    shp = L.make_shape(B, G, F, H, W)
    E.call("glrgtv_block_fwd", shp, p, x, f0.detach(), f1.detach(), out, sv, None)
    gr, gkeep = alloc_grads(G, F)
    nbytes = E.emu_lib().glrgtv_block_bwd_workspace_bytes(shp)
    ws = torch.empty(nbytes // 4)
    gx, gf0, gf1 = torch.empty_like(x), torch.empty_like(f0), torch.empty_like(f1)
    E.call("glrgtv_block_bwd", shp, p, x, f0.detach(), f1.detach(), sv, gout, gx, gf0, gf1, gr, ws, nbytes, None)
    names = list(pw)
    gfeat = torch.autograd.grad([f0, f1], [xx] + [pw[k] for k in names], [gf0, gf1])
    gx_total = gx + gfeat[0]
    assert rel(gx_total, gx_ref) < 5e-5, rel(gx_total, gx_ref)
    got = grads_to_state_names(gkeep, G, F)
    got.update({k: g for k, g in zip(names, gfeat[1:])})
    for k, ref in pg_ref.items():
        if float(ref.abs().max()) == 0.0:
            assert float(got[k].abs().max()) == 0.0, k
        else:
            assert rel(got[k], ref) < 2e-4, (k, rel(got[k], ref))
