"""LocalNonLinearBlock inference forward (host_cnn.py over glrgtv_pixel_rstd / glrgtv_dwconv_gate) on the g++ EMULATION build,
against the module itself (the reference's op sequence, V1X0:911-964, in plain PyTorch)."""
import pytest
import torch

from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
from imagerestoration_development_unrolling_b200 import host_cnn
from tests import emu_harness as E

pytestmark = pytest.mark.usefixtures("isolated_rng")


class EmuCnnKernels:
    def pixel_rstd(self, x, nsub, eps):
        B, C, H, W = x.shape
        rs = torch.full((B, nsub, H, W), float("nan"))
        E.call("glrgtv_pixel_rstd", B, C, nsub, H * W, float(eps), x, rs, None)
        return rs

    def dwconv_gate(self, h, rs, w9, top, bot):
        B, C2, H, W = h.shape
        u = torch.full((B, C2 // 2, H, W), float("nan"))
        E.call("glrgtv_dwconv_gate", B, C2 // 2, rs.shape[1], H, W, h.contiguous(), rs, w9, top, bot, u, None)
        return u


    def dwconv_gate_bwd(self, h, rs, w9, gu):
        B, C2, H, W = h.shape
        gM, gh, gw9 = torch.full_like(h, float("nan")), torch.full_like(h, float("nan")), torch.zeros_like(w9)
        E.call("glrgtv_dwconv_gate_bwd", B, C2 // 2, rs.shape[1], H, W, h.contiguous(), rs, w9, gu.contiguous(), gM, gh, gw9, None)
        return gh, gw9

    def pixel_norm_bwd(self, x, rs, gx1, gout, s0, nsub):
        B, C, H, W = x.shape
        gx = torch.full_like(x, float("nan"))
        E.call("glrgtv_pixel_norm_bwd", B, C, nsub, H * W, x, rs, gx1.contiguous(), gout.contiguous(), s0.contiguous(), gx, None)
        return gx


def _block(dim, hidden, nsub, seed):
    torch.manual_seed(seed)
    blk = M.LocalNonLinearBlock(dim, hidden, nsub).eval()
    with torch.no_grad():
        blk.norm.weighted_transform.weight.uniform_(0.5, 1.5)
        blk.skip_weight.copy_(torch.tensor([0.9, 0.7]))
    return blk


@pytest.mark.parametrize("dim,hidden,nsub,B,H,W", [
    (8, 12, 1, 2, 9, 12),      # odd height, one band
    (12, 8, 2, 1, 70, 20),     # two sub-nets (gate and value halves scale with different sub-nets), three row bands
    (6, 4, 1, 1, 1, 4),        # a single row, a single quad: every tap replicates
    (16, 16, 4, 1, 33, 8),     # band boundary at row 32
    (4, 2, 1, 1, 3, 260),      # 65 quads per row: warps straddle rows, lanes 0 / 31 load their edge scalars themselves
    (4, 4, 2, 2, 34, 132),     # the same with two sub-nets and a short last band
])
def test_nonlinear_block_matches_module(dim, hidden, nsub, B, H, W):
    blk = _block(dim, hidden, nsub, seed=H)
    x = torch.randn(B, dim, H, W, generator=torch.Generator().manual_seed(W)) * 2 + 0.5
    with torch.no_grad():
        ref = blk(x)
    got = host_cnn.nonlinear_block_forward(blk, x, EmuCnnKernels())
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert float((got - ref).abs().max()) < 2e-5 * float(ref.abs().max())


def test_pixel_rstd_large_mean():
    """features with a mean far above their spread: the variance must neither cancel (sum / sum-of-squares) nor pick up the rounding
    of a running mean to first order (one-pass Welford: 1e-5 here) - the kernel is two-pass like torch.var"""
    x = 100.0 + 0.01 * torch.randn(1, 48, 4, 8, generator=torch.Generator().manual_seed(1))
    rs = EmuCnnKernels().pixel_rstd(x, 1, 1e-5)
    ref = 1.0 / torch.sqrt(x.double().var(dim=1, keepdim=True, correction=1) + 1e-5)
    assert float(((rs - ref) / ref).abs().max()) < 2e-6


@pytest.mark.parametrize("c", [2, 3, 48])
def test_pixel_rstd_matches_torch_accuracy(c):
    """as accurate as torch's own fp32 variance, also for two or three channels per sub-net where nearly equal channel values make
    rs large (found by a random-shape sweep: a Welford update was 100x less accurate than torch there)"""
    x = torch.randn(2, c, 33, 64, generator=torch.Generator().manual_seed(c)) * 2 + 0.5
    rs = EmuCnnKernels().pixel_rstd(x, 1, 1e-5)
    ref64 = 1.0 / torch.sqrt(x.double().var(dim=1, keepdim=True, correction=1) + 1e-5)
    ref32 = 1.0 / torch.sqrt(x.var(dim=1, keepdim=True, correction=1) + 1e-5)
    err = float(((rs.double() - ref64) / ref64).abs().max())
    assert err < 4 * max(float(((ref32.double() - ref64) / ref64).abs().max()), 1e-7), err


def test_strip_rows_from_neighbours():
    """a 3-strip split with the neighbours' scaled rows handed in as top / bot equals the whole image"""
    blk = _block(8, 8, 1, seed=5)
    x = torch.randn(1, 8, 48, 16, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = blk(x)
    K = EmuCnnKernels()
    bounds = [(0, 16), (16, 35), (35, 48)]
    firsts, lasts = {}, {}

    def record(i):
        def ex(first, last):
            firsts[i], lasts[i] = first, last
            return None, None
        return ex

    for i, (a, b) in enumerate(bounds):                      # pass 1: collect every strip's scaled border rows
        host_cnn.nonlinear_block_forward(blk, x[:, :, a:b], K, record(i))
    outs = []
    for i, (a, b) in enumerate(bounds):                      # pass 2: hand the neighbours' rows in
        ex = lambda f, l, i=i: (lasts.get(i - 1), firsts.get(i + 1))     # noqa: E731
        outs.append(host_cnn.nonlinear_block_forward(blk, x[:, :, a:b], K, ex))
    got = torch.cat(outs, 2)
    assert float((got - ref).abs().max()) < 2e-5 * float(ref.abs().max())


def test_argument_checks():
    x = torch.zeros(1, 4, 2, 6)
    with pytest.raises(RuntimeError, match="unsupported|UNSUPPORTED|-6"):
        E.call("glrgtv_pixel_rstd", 1, 4, 1, 2 * 3, 1e-5, x, x, None)             # HW % 4 != 0
    with pytest.raises(RuntimeError):
        E.call("glrgtv_pixel_rstd", 1, 4, 4, 8, 1e-5, x, x, None)                 # one channel per sub-net: no variance
    with pytest.raises(RuntimeError):
        E.call("glrgtv_dwconv_gate", 1, 2, 1, 2, 6, x, x, x, None, None, x, None)  # W % 4 != 0


@pytest.mark.parametrize("dim,hidden,nsub,B,H,W", [
    (8, 12, 1, 2, 9, 12),
    (12, 8, 2, 1, 70, 20),
    (6, 4, 1, 1, 1, 4),        # one row, one quad: every padded tap folds back onto the pixel itself
    (6, 4, 1, 1, 2, 8),
    (16, 16, 4, 1, 33, 8),
    (4, 2, 1, 1, 3, 260),
])
def test_nonlinear_block_gradients_match_autograd(dim, hidden, nsub, B, H, W):
    """input and all six parameter gradients of the kernel path against autograd through the module (double precision)"""
    blk = _block(dim, hidden, nsub, seed=H + 1)
    x = (torch.randn(B, dim, H, W, generator=torch.Generator().manual_seed(W + 1)) * 2 + 0.5).requires_grad_(True)
    gout = torch.randn(B, dim, H, W, generator=torch.Generator().manual_seed(3))
    params = [blk.norm.weighted_transform.weight, blk.local_linear.channels_linear_op.weight, blk.local_linear.channels_local_linear_op.weight,
              blk.local_linear.project_out.weight, blk.skip_weight]
    out = host_cnn.nonlinear_block_train(blk, x, EmuCnnKernels())
    got = torch.autograd.grad(out, [x] + params, gout)
    import copy
    ref_blk = copy.deepcopy(blk).double()
    xd = x.detach().double().requires_grad_(True)
    rp = [ref_blk.norm.weighted_transform.weight, ref_blk.local_linear.channels_linear_op.weight, ref_blk.local_linear.channels_local_linear_op.weight,
          ref_blk.local_linear.project_out.weight, ref_blk.skip_weight]
    ref_out = ref_blk(xd)
    ref = torch.autograd.grad(ref_out, [xd] + rp, gout.double())
    assert float((out.detach().double() - ref_out.detach()).abs().max()) < 2e-5 * float(ref_out.detach().abs().max())
    for name, g, r in zip(["x", "norm", "linear", "depthwise", "project_out", "skip"], got, ref):
        assert g.shape == r.shape, name
        assert torch.isfinite(g).all(), name
        assert float((g.double() - r).norm()) < 2e-5 * float(r.norm()) + 1e-9, name


def test_switch_is_a_noop_for_cpu_tensors():
    """the switch (on by default) only reroutes CUDA float32 inputs; CPU tensors keep the PyTorch op sequence whatever it says (the
    host CNN is not the library's hot path, so it may run anywhere the reference's does)"""
    blk = _block(8, 8, 1, seed=2)
    x = torch.randn(1, 8, 6, 8)
    with torch.no_grad():
        prev = M.set_host_cnn_kernels(False)
        try:
            ref = blk(x)
            M.set_host_cnn_kernels(True)
            got = blk(x)
        finally:
            M.set_host_cnn_kernels(prev)
    assert prev is True and torch.equal(got, ref)
