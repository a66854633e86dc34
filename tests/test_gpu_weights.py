"""Edge-weight row walkers (csrc/weights_walk.cu) on the GPU against the round-1 tile kernels (the second implementation) at the
benchmark's plane sizes and on a 4K-wide band (strips with seams), and against the oracle through the whole block."""
import pytest
import torch

from oracle import glr_gtv_oracle as O
from tests.util import rel, random_block_state
from tests.test_gpu_block import make_block, run_block, check_against

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as m
    return m


@pytest.fixture()
def lib():
    from imagerestoration_development_unrolling_b200 import _lib as L
    lib = L.load()
    yield lib
    lib.glrgtv_set_weights_kernels(0)


@pytest.mark.parametrize("scale", [0, 1, 2, 3])
def test_walkers_equal_tile_kernels_at_benchmark_size(M, lib, scale):
    dim, G = [48, 96, 192, 384][scale], [8, 16, 16, 32][scale]
    B, H = 2, 256 >> scale
    sd = random_block_state(dim, G, seed=61 + scale)
    gen = torch.Generator().manual_seed(70 + scale)
    x, gout = torch.randn(B, dim, H, H, generator=gen), torch.randn(B, dim, H, H, generator=gen)
    blk = make_block(M, dim, G, sd)
    lib.glrgtv_set_weights_kernels(1)
    out1, gx1, pg1 = run_block(blk, x, gout)
    lib.glrgtv_set_weights_kernels(0)
    n0 = lib.glrgtv_weights_walk_launch_count()
    out0, gx0, pg0 = run_block(blk, x, gout)
    assert lib.glrgtv_weights_walk_launch_count() - n0 == 4          # forward + backward, full and half resolution
    assert rel(out0, out1) < 2e-6, rel(out0, out1)
    assert rel(gx0, gx1) < 2e-5, rel(gx0, gx1)
    for k in pg1:
        if float(pg1[k].abs().max()) == 0.0:
            assert float(pg0[k].abs().max()) == 0.0, k
        else:
            # (threshold-crossing elements flip with the rounding of the weights: see test_gpu_stream.py)
            tol = 5e-3 if "gamma" in k else 1e-3
            assert rel(pg0[k], pg1[k]) < tol, (k, rel(pg0[k], pg1[k]))


@pytest.mark.parametrize("case", [(12, 2, 2, 24, 264), (24, 2, 1, 40, 136), (12, 2, 2, 70, 40), (48, 8, 1, 64, 64)])
def test_walkers_against_oracle(M, lib, case):
    """seams at both resolutions (F = 6: 66 / 33 quads; F = 12: 68 / 34 pairs), ragged walkers, row bands with halo rows"""
    dim, G, B, H, W = case
    sd = random_block_state(dim, G, seed=dim + H + W)
    gen = torch.Generator().manual_seed(H + 3 * W)
    x, gout = torch.randn(B, dim, H, W, generator=gen), torch.randn(B, dim, H, W, generator=gen)
    ref = O.lowpass_block_fwd_bwd({k: v.double() for k, v in sd.items()}, x.double(), gout.double())
    n0 = lib.glrgtv_weights_walk_launch_count()
    out, gx, pg = run_block(make_block(M, dim, G, sd), x, gout)
    assert lib.glrgtv_weights_walk_launch_count() - n0 == 4
    check_against(out, gx, pg, *ref)


def test_4k_band_forward(M, lib):
    """a 3840-wide band (30 strips of 128 columns, 15 at half resolution): walkers against the tile kernels"""
    dim, G = 48, 8
    blk = make_block(M, dim, G, random_block_state(dim, G, seed=2))
    x = torch.randn(1, dim, 48, 3840, generator=torch.Generator().manual_seed(4)).cuda()
    with torch.no_grad():
        lib.glrgtv_set_weights_kernels(1)
        ref = blk(x)
        lib.glrgtv_set_weights_kernels(0)
        out = blk(x)
    assert rel(out, ref) < 2e-6, rel(out, ref)
