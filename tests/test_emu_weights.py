"""Edge-weight row walkers (csrc/weights_walk.cu) in the g++ emulation build (fibers: real warp shuffles) against the oracle and
against the round-1 tile kernels they replace - forward weights of both levels and both families, the feature gradients and the
multiM gradients, through glrgtv_block_fwd / glrgtv_block_bwd."""
import pytest
import torch

from oracle import glr_gtv_oracle as O
from imagerestoration_development_unrolling_b200 import _lib as L
from tests import emu_harness as E
from tests.util import rel, random_block_state, block_structs, alloc_saved, oracle_features, alloc_grads, grads_to_state_names

# (dim, ngraphs, B, H, W), F = dim / ngraphs = 6 or 12.  Several walkers per warp (W = 8 .. 64), ragged walkers (W = 40, 24: lanes
# beyond the width), strips with seams at both resolutions (F = 6: W = 264 -> 66 quads = 3 strips, half resolution 33 quads = 2
# strips; F = 12: W = 136 -> 68 pairs = 3 strips), row bands with halo rows (H = 40, 72), both families, batch > 1
CASES = [(12, 2, 2, 12, 16), (6, 1, 1, 10, 8), (12, 2, 1, 40, 40), (24, 2, 1, 6, 24), (6, 1, 1, 8, 264), (12, 1, 1, 6, 136),
         (12, 2, 1, 72, 32), (24, 2, 2, 4, 64)]


def run(case, generation):
    dim, G, B, H, W = case
    F = dim // G
    lib = E.emu_lib()
    lib.glrgtv_set_weights_kernels(generation)
    try:
        sd = random_block_state(dim, G, seed=700 + H + W)
        gen = torch.Generator().manual_seed(H * W + 7)
        x, gout = torch.randn(B, dim, H, W, generator=gen), torch.randn(B, dim, H, W, generator=gen)
        f0, f1 = oracle_features(sd, x)
        p, keep = block_structs(sd)
        sv, saved = alloc_saved(B, G, F, H, W)
        out = torch.empty_like(x)
        shp = L.make_shape(B, G, F, H, W)
        n0 = lib.glrgtv_weights_walk_launch_count()
        E.call("glrgtv_block_fwd", shp, p, x, f0, f1, out, sv, None)
        gr, gkeep = alloc_grads(G, F)
        nbytes = lib.glrgtv_block_bwd_workspace_bytes(shp)
        ws = torch.empty(nbytes // 4)
        gx, gf0, gf1 = torch.empty_like(x), torch.empty_like(f0), torch.empty_like(f1)
        E.call("glrgtv_block_bwd", shp, p, x, f0, f1, sv, gout, gx, gf0, gf1, gr, ws, nbytes, None)
        got = grads_to_state_names(gkeep, G, F)
        # two forward + two backward walker launches (full and half resolution), or none with the tile kernels
        assert lib.glrgtv_weights_walk_launch_count() - n0 == (4 if generation == 0 else 0)
    finally:
        lib.glrgtv_set_weights_kernels(0)
    names = {k: saved[k] for k in ("wT0", "wL0", "wT1", "wL1")}
    return sd, x, gout, names, gf0, gf1, {k: v for k, v in got.items() if "multiM" in k}


@pytest.mark.parametrize("case", CASES)
def test_walkers_equal_tile_kernels_and_oracle(case):
    dim, G, B, H, W = case
    sd, x, gout, w_new, gf0_new, gf1_new, gM_new = run(case, 0)
    _, _, _, w_old, gf0_old, gf1_old, gM_old = run(case, 1)
    for k in w_new:                                   # same weights (fp32 rounding apart) from both implementations
        assert rel(w_new[k], w_old[k]) < 2e-6, (k, rel(w_new[k], w_old[k]))
        assert float((w_new[k].sum(dim=2) - 1).abs().max()) < 1e-5          # a softmax over the four edges
    assert rel(gf0_new, gf0_old) < 2e-5, rel(gf0_new, gf0_old)
    assert rel(gf1_new, gf1_old) < 2e-5, rel(gf1_new, gf1_old)
    for k in gM_new:
        assert rel(gM_new[k], gM_old[k]) < 2e-5, (k, rel(gM_new[k], gM_old[k]))
    # and the oracle: weights through the forward intermediates, multiM gradients through autograd
    _, inter = O.mixture_gtvglr_forward({k: v.double() for k, v in sd.items()}, x.double(), "local_filter.", return_intermediates=True)
    for k in w_new:
        if k in inter:
            assert rel(w_new[k], inter[k].reshape(w_new[k].shape)) < 1e-5, k
    _, _, pg_ref = O.lowpass_block_fwd_bwd({k: v.double() for k, v in sd.items()}, x.double(), gout.double())
    for k in gM_new:
        assert rel(gM_new[k], pg_ref[k]) < 2e-4, (k, rel(gM_new[k], pg_ref[k]))


def test_other_shapes_stay_on_the_tile_kernels():
    """F = 3 is not a walker shape; an unknown generation is refused"""
    lib = E.emu_lib()
    assert lib.glrgtv_set_weights_kernels(2) != 0
    dim, G, B, H, W = 3, 1, 1, 8, 16
    sd = random_block_state(dim, G, seed=1)
    x = torch.randn(B, dim, H, W, generator=torch.Generator().manual_seed(2))
    f0, f1 = oracle_features(sd, x)
    p, keep = block_structs(sd)
    sv, saved = alloc_saved(B, G, 3, H, W)
    n0 = lib.glrgtv_weights_walk_launch_count()
    E.call("glrgtv_block_fwd", L.make_shape(B, G, 3, H, W), p, x, f0, f1, torch.empty_like(x), sv, None)
    assert lib.glrgtv_weights_walk_launch_count() == n0
