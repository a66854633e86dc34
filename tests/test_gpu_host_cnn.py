"""LocalNonLinearBlock inference forward on libglrgtv's kernels (host_cnn.py: glrgtv_pixel_rstd, glrgtv_dwconv_gate + cuBLAS
GEMMs) against the module's own op sequence (V1X0:911-964) on the GPU, and the whole network through it against the
golden produced by the reference's AbtractMultiScaleGraphFilter."""
import os

import numpy as np
import pytest
import torch

from tests.util import rel

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("isolated_rng")]


def _block(dim, hidden, nsub, seed):
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
    torch.manual_seed(seed)
    blk = M.LocalNonLinearBlock(dim, hidden, nsub).eval()
    with torch.no_grad():
        blk.norm.weighted_transform.weight.uniform_(0.5, 1.5)
        blk.skip_weight.copy_(torch.tensor([0.9, 0.7]))
    return blk.cuda()


@pytest.mark.parametrize("dim,hidden,nsub,B,H,W", [(48, 96, 1, 2, 64, 64), (96, 192, 1, 1, 70, 52), (24, 16, 2, 3, 33, 8), (384, 768, 1, 1, 16, 32)])
def test_nonlinear_block_matches_module(dim, hidden, nsub, B, H, W):
    from imagerestoration_development_unrolling_b200 import host_cnn
    blk = _block(dim, hidden, nsub, seed=H)
    x = torch.randn(B, dim, H, W, generator=torch.Generator().manual_seed(W)).cuda() * 2 + 0.5
    with torch.no_grad():
        ref = blk(x)
    got = host_cnn.nonlinear_block_forward(blk, x)
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert rel(got, ref) < 1e-5


def test_strip_rows_from_neighbours():
    from imagerestoration_development_unrolling_b200 import host_cnn
    blk = _block(48, 96, 1, seed=5)
    x = torch.randn(1, 48, 96, 64, generator=torch.Generator().manual_seed(2)).cuda()
    with torch.no_grad():
        ref = blk(x)
    bounds, firsts, lasts = [(0, 32), (32, 70), (70, 96)], {}, {}
    for i, (a, b) in enumerate(bounds):
        def record(first, last, i=i):
            firsts[i], lasts[i] = first, last
            return None, None
        host_cnn.nonlinear_block_forward(blk, x[:, :, a:b], None, record)
    outs = [host_cnn.nonlinear_block_forward(blk, x[:, :, a:b], None, lambda f, l, i=i: (lasts.get(i - 1), firsts.get(i + 1)))
            for i, (a, b) in enumerate(bounds)]
    assert rel(torch.cat(outs, 2), ref) < 1e-5


def test_whole_network_through_the_fused_cnn_matches_reference_golden(golden_dir):
    """the reference's own output (tests/golden/model_small.npz) reproduced with every LocalNonLinearBlock on the kernels"""
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M, shard
    from tests.test_gpu_model import CFG
    z = np.load(os.path.join(golden_dir, "model_small.npz"))
    m = M.AbtractMultiScaleGraphFilter(**CFG)
    m.load_state_dict({str(k): torch.from_numpy(z["sd." + str(k)]) for k in z["keys"]}, strict=True)
    m = m.cuda().eval()
    noisy = torch.from_numpy(z["noisy"]).float().cuda()
    ex = shard.ShardedMultiScaleFilter(m, 0, 1)
    out = ex(noisy)
    assert rel(out, torch.from_numpy(z["out"])) < 1e-4
    with torch.no_grad():
        assert rel(ex.enc_dec(noisy), m.enc_dec(noisy)) < 1e-5


def test_cpu_tensor_raises():
    from imagerestoration_development_unrolling_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.pixel_rstd(torch.zeros(1, 4, 2, 4), 1, 1e-5)
