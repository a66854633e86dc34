"""LocalNonLinearBlock forward + backward on libglrgtv's kernels (host_cnn.nonlinear_block_train, opt-in) against autograd through
the module (double precision) on the GPU.  The same comparison runs on CPU on the emulation build (tests/test_emu_host_cnn.py)."""
import copy

import pytest
import torch

from tests.util import rel

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("isolated_rng")]


@pytest.mark.parametrize("dim,hidden,nsub,B,H,W", [(48, 96, 1, 2, 64, 64), (24, 16, 2, 3, 33, 8), (96, 192, 1, 1, 70, 52)])
def test_nonlinear_block_gradients_match_autograd(dim, hidden, nsub, B, H, W):
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M, host_cnn
    torch.manual_seed(H)
    blk = M.LocalNonLinearBlock(dim, hidden, nsub)
    with torch.no_grad():
        blk.norm.weighted_transform.weight.uniform_(0.5, 1.5)
        blk.skip_weight.copy_(torch.tensor([0.9, 0.7]))
    blk = blk.cuda()
    ref_blk = copy.deepcopy(blk).double()
    x = (torch.randn(B, dim, H, W, generator=torch.Generator().manual_seed(W)) * 2 + 0.5).cuda().requires_grad_(True)
    gout = torch.randn(B, dim, H, W, generator=torch.Generator().manual_seed(3)).cuda()

    def params(b):
        return [b.norm.weighted_transform.weight, b.local_linear.channels_linear_op.weight, b.local_linear.channels_local_linear_op.weight,
                b.local_linear.project_out.weight, b.skip_weight]

    out = host_cnn.nonlinear_block_train(blk, x)
    got = torch.autograd.grad(out, [x] + params(blk), gout)
    xd = x.detach().double().requires_grad_(True)
    ref_out = ref_blk(xd)
    ref = torch.autograd.grad(ref_out, [xd] + params(ref_blk), gout.double())
    assert rel(out.detach(), ref_out.detach()) < 1e-5
    for name, g, r in zip(["x", "norm", "linear", "depthwise", "project_out", "skip"], got, ref):
        assert g.shape == r.shape and torch.isfinite(g).all(), name
        assert rel(g, r) < 1e-4, name
