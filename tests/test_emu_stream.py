"""Register-streaming block kernels (csrc/block_stream_*.cu) in the g++ emulation build (fibers give the warp
shuffles and barriers their real meaning) against the oracle.  The path is forced with glrgtv_set_block_path(2)."""
import pytest
import torch

from oracle import glr_gtv_oracle as O
from imagerestoration_development_unrolling_b200 import _lib as L
from tests import emu_harness as E
from tests.util import (rel, random_block_state, block_structs, alloc_saved, oracle_features, alloc_grads,
                        grads_to_state_names)

# (dim, ngraphs, B, H, W): every walker width (8/16/32/64 lanes), partial walkers (W=24, 48, 136), bands, tiny planes
# the tall last case makes the planner cut the plane into row BANDS (few CTAs otherwise): walkers that start mid-image
CASES = [(12, 2, 2, 12, 16), (6, 1, 1, 10, 8), (12, 2, 1, 20, 24), (12, 4, 1, 14, 40), (6, 2, 1, 8, 64),
         (6, 1, 1, 12, 72), (6, 2, 1, 6, 128), (3, 1, 1, 8, 136), (2, 1, 1, 6, 256), (24, 2, 1, 2, 8), (3, 1, 1, 128, 16)]


@pytest.fixture(autouse=True, params=[1, 2], ids=["cp_async", "tma"])
def stream_path(request):
    E.emu_lib().glrgtv_set_block_path(2)
    E.emu_lib().glrgtv_set_stream_loader(request.param)
    yield
    E.emu_lib().glrgtv_set_block_path(0)
    E.emu_lib().glrgtv_set_stream_loader(0)


# planes wider than one 64-lane walker are cut into column strips (forward only): 2 and 3 strips, ragged last strip
WIDE = [(6, 1, 1, 8, 272), (4, 2, 1, 6, 504), (2, 1, 1, 4, 488), (6, 1, 1, 6, 504), (12, 1, 1, 4, 520)]


def fw2_eligible(F, H, W):
    """mirror of glr_fw2_eligible (csrc/fw2.cu): W % 8 == 0; a CTA of <= 384 threads holds at least half of a graph's channels on a
    window of >= 16 columns at both resolutions (wider planes: column strips)"""
    if W % 8 or H % 2 or W < 16 or H < 4:
        return False
    for lw in (W, W // 2):
        lanes = 4
        while 2 * lanes < lw:
            lanes *= 2
        nch = lambda L: max([n for n in range(1, F + 1) if F % n == 0 and n * L <= 384], default=0)
        while lanes > 4 and nch(lanes) * 2 < F:
            lanes //= 2
        if lanes < 8 or nch(lanes) < 1 or 6 * nch(lanes) < 10:
            return False
    return True


@pytest.mark.parametrize("generation", [1, 2], ids=["round1", "fw2"])
@pytest.mark.parametrize("case", CASES + WIDE)
def test_stream_block_forward(case, generation):
    dim, G, B, H, W = case
    F = dim // G
    E.emu_lib().glrgtv_set_fwd_kernels(generation)
    sd = random_block_state(dim, G, seed=300 + H)
    x = torch.randn(B, dim, H, W, generator=torch.Generator().manual_seed(H * W))
    ref_out, inter = O.mixture_gtvglr_forward({k: v.double() for k, v in sd.items()}, x.double(), "local_filter.",
                                              return_intermediates=True)
    s = sd["skip_weight"].double()
    ref_out = s[0] * x.double() + s[1] * ref_out
    f0, f1 = oracle_features(sd, x)
    p, keep = block_structs(sd)
    sv, saved = alloc_saved(B, G, F, H, W)
    out = torch.empty_like(x)
    n0 = E.emu_lib().glrgtv_stream_launch_count()
    E.call("glrgtv_block_fwd", L.make_shape(B, G, F, H, W), p, x, f0, f1, out, sv, None)
    nl = E.emu_lib().glrgtv_stream_launch_count() - n0
    E.emu_lib().glrgtv_set_fwd_kernels(0)
    # round 1: one launch per stage; fw2: a half-resolution and a full-resolution launch per stage
    assert nl == (8 if generation == 2 and fw2_eligible(F, H, W) else 4), nl
    for n in ("wT0", "wL0", "wT1", "wL1"):
        assert rel(saved[n], inter[n]) < 5e-6, n
    for n in ("bA", "x1", "bB", "r1", "x2"):
        assert rel(saved[n], inter[n]) < 2e-5, (n, rel(saved[n], inter[n]))
    assert rel(out, ref_out) < 1e-5


# second-generation pair walkers (csrc/bw2.cu): extra shapes - wide walkers with seams (W = 256: four warps per row), ragged
# widths, several channels per warp, row bands, F = 12
BW2_CASES = CASES + [(3, 1, 1, 6, 272), (2, 1, 1, 4, 504), (6, 1, 1, 8, 256), (12, 1, 1, 6, 64), (6, 2, 2, 60, 8), (3, 1, 1, 100, 24),
                     (12, 1, 1, 6, 128), (6, 1, 2, 10, 136), (24, 2, 1, 6, 32)]


def bw2_eligible(F, H, W):
    """mirror of glr_bw2_eligible (csrc/bw2.cu): W % 8 == 0, a CTA of <= 384 threads holds at least half of a graph's channels
    (>= 2 of them: two weight-plane copies per thread) at both resolutions"""
    if W % 8 or H % 2 or W < 8 or H < 4:
        return False
    for lw in (W, W // 2):
        lanes = 4
        while 2 * lanes < lw:
            lanes *= 2
        nch = max([n for n in range(1, F + 1) if F % n == 0 and n * lanes <= 384], default=0)
        if nch < 2 or F // nch > 2:
            return False
    return True


@pytest.mark.parametrize("generation", [1, 2], ids=["round1", "bw2"])
@pytest.mark.parametrize("case", BW2_CASES)
def test_stream_block_backward(case, generation):
    dim, G, B, H, W = case
    bw2 = generation == 2
    if not bw2 and case not in CASES + [(3, 1, 1, 6, 272), (2, 1, 1, 4, 504)]:
        pytest.skip("round-1 kernels: covered by their own cases")
    E.emu_lib().glrgtv_set_bwd_kernels(generation)
    F = dim // G
    sd = random_block_state(dim, G, seed=400 + H)
    gen = torch.Generator().manual_seed(H * W + 1)
    x = torch.randn(B, dim, H, W, generator=gen)
    gout = torch.randn(B, dim, H, W, generator=gen)
    sd64 = {k: v.double() for k, v in sd.items()}
    _, gx_ref, pg_ref = O.lowpass_block_fwd_bwd(sd64, x.double(), gout.double())
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
    n0 = E.emu_lib().glrgtv_stream_launch_count()
    E.call("glrgtv_block_bwd", shp, p, x, f0.detach(), f1.detach(), sv, gout, gx, gf0, gf1, gr, ws, nbytes, None)
    E.emu_lib().glrgtv_set_bwd_kernels(2)
    nl = E.emu_lib().glrgtv_stream_launch_count() - n0
    # round 1: 5 stage kernels + 4 x 2 edge-weight-gradient kernels; bw2: 5 stages x (half + full resolution), no gradient pass
    assert nl == (10 if bw2 and bw2_eligible(F, H, W) else 13), nl
    names = list(pw)
    gfeat = torch.autograd.grad([f0, f1], [xx] + [pw[k] for k in names], [gf0, gf1])
    gx_total = gx + gfeat[0]
    got = grads_to_state_names(gkeep, G, F)
    got.update({k: g for k, g in zip(names, gfeat[1:])})
    errs = {k: (rel(got[k], ref) if float(ref.abs().max()) > 0 else float(got[k].abs().max())) for k, ref in pg_ref.items()}
    errs["gx"] = rel(gx_total, gx_ref)
    bad = {k: v for k, v in errs.items() if v > 2e-4}
    assert not bad, bad


@pytest.mark.parametrize("case", [(12, 2, 1, 40, 16), (6, 1, 2, 64, 24), (4, 2, 1, 24, 272)])
def test_stage_entry_point_on_row_ranges(case):
    """glrgtv_block_fwd_stage: the block forward as weights + four stages, every stage run on three row ranges in turn
    (what a spatially sharded caller does between halo exchanges) equals the oracle"""
    dim, G, B, H, W = case
    F = dim // G
    sd = random_block_state(dim, G, seed=500 + H)
    x = torch.randn(B, dim, H, W, generator=torch.Generator().manual_seed(H + W))
    ref_out, inter = O.mixture_gtvglr_forward({k: v.double() for k, v in sd.items()}, x.double(), "local_filter.",
                                              return_intermediates=True)
    s = sd["skip_weight"].double()
    ref_out = s[0] * x.double() + s[1] * ref_out
    f0, f1 = oracle_features(sd, x)
    p, keep = block_structs(sd)
    sv, saved = alloc_saved(B, G, F, H, W)
    for t in saved.values():
        t.fill_(float("nan"))                      # a stage that reads rows nobody produced shows up as NaN
    out = torch.full_like(x, float("nan"))
    shp = L.make_shape(B, G, F, H, W)
    E.call("glrgtv_block_fwd_stage", 0, shp, p, x, f0, f1, out, sv, 0, H, None)
    cuts = [0, (H // 3) & ~1, (2 * H // 3) & ~1, H]
    for stage in (1, 2, 3, 4):
        for a, b in zip(cuts[:-1], cuts[1:]):
            E.call("glrgtv_block_fwd_stage", stage, shp, p, x, None, None, out, sv, a, b, None)
    for n in ("bA", "x1", "bB", "r1", "x2"):
        assert rel(saved[n], inter[n]) < 2e-5, (n, rel(saved[n], inter[n]))
    assert rel(out, ref_out) < 1e-5
