"""Per-operator kernels (csrc/ops_basic.cu), compiled by g++ in emulation mode, against the oracle and
its autograd.  This is a CPU check of the kernel SOURCES' arithmetic/border logic; the same assertions
run against the real nvcc build on the GPU in test_gpu_ops.py."""
import pytest
import torch

from oracle import glr_gtv_oracle as O
from imagerestoration_development_unrolling_b200 import _lib as L
from tests import emu_harness as E

WINDOWS = ["cross3", "full3", "small5"]
SHAPES = [(2, 2, 3, 5, 7), (1, 3, 6, 2, 2), (1, 1, 4, 1, 6), (1, 2, 2, 6, 1), (1, 1, 3, 4, 4)]


@pytest.fixture(autouse=True)
def _seed(request):
    """every test draws from its own seeded stream: the data does not depend on which tests ran before"""
    import zlib
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(zlib.crc32(request.node.nodeid.encode()))
        yield


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def rnd(*s):
    return torch.randn(*s, dtype=torch.float32)


def stats_params(n):
    return [torch.full((n, 1, 1, 1), v) + 0.2 * rnd(n, 1, 1, 1) for v in (1.0, 0.5, 0.5, 0.5)]


@pytest.mark.parametrize("window", WINDOWS)
@pytest.mark.parametrize("shape", SHAPES)
def test_edge_weights(window, shape):
    B, G, F, H, W = shape
    edges = O.window_edges(window)
    Ne = len(edges)
    feat = rnd(B, G, F, H, W).requires_grad_(True)
    M = (1 + 0.5 * rnd(G, F)).requires_grad_(True)
    w_ref = O.edge_weights(feat, M, edges)
    gw = rnd(B, G, Ne, H, W)
    gfeat_ref, gM_ref = torch.autograd.grad(w_ref, [feat, M], gw)
    shp, win = L.make_shape(*shape), L.make_window(edges)
    w = torch.empty(B, G, Ne, H, W)
    E.call("glrgtv_edge_weights_fwd", shp, win, feat.detach(), M.detach(), w, None)
    assert rel(w, w_ref) < 2e-6
    gfeat, gM = torch.empty_like(feat), torch.zeros_like(M)
    scratch = torch.empty(B * G * (Ne + 1) * H * W)
    E.call("glrgtv_edge_weights_bwd", shp, win, feat.detach(), M.detach(), w, gw, gfeat, gM, scratch, None)
    assert rel(gfeat, gfeat_ref) < 2e-5
    assert rel(gM, gM_ref) < 2e-5


@pytest.mark.parametrize("pad", ["clamp", "reflect"])
@pytest.mark.parametrize("per_channel", [True, False])
@pytest.mark.parametrize("shape", SHAPES)
def test_stats_conv_and_transpose(pad, per_channel, shape):
    B, G, F, H, W = shape
    if pad == "reflect" and (H < 2 or W < 2):
        pytest.skip("reflect needs >= 2 pixels")
    n = G * F if per_channel else 1
    ps = [p.requires_grad_(True) for p in stats_params(n)]
    x = rnd(B, G, F, H, W).requires_grad_(True)
    g = rnd(B, G, F, H, W)
    shp = L.make_shape(*shape)
    st = L.make_stats(*[p.detach() for p in ps], pad=L.PAD_CLAMP if pad == "clamp" else L.PAD_REFLECT)
    for fwd, bwd, ref_fn in (("glrgtv_stats_conv_fwd", "glrgtv_stats_conv_bwd", lambda: O.stats_conv(x, ps, pad)),
                             ("glrgtv_stats_conv_t_fwd", "glrgtv_stats_conv_t_bwd", lambda: O.stats_conv_transpose(x, ps))):
        ref = ref_fn()
        grads = torch.autograd.grad(ref, [x] + ps, g)
        out = torch.empty_like(x)
        E.call(fwd, shp, st, x.detach(), out, None)
        assert rel(out, ref) < 1e-6
        gx, gst = torch.empty_like(x), torch.zeros(4 * n)
        E.call(bwd, shp, st, x.detach(), g, gx, gst, None)
        assert rel(gx, grads[0]) < 1e-6
        for i in range(4):
            assert rel(gst[i * n:(i + 1) * n], grads[1 + i].reshape(-1)) < 2e-5


@pytest.mark.parametrize("window", WINDOWS)
@pytest.mark.parametrize("shape", SHAPES)
def test_L_C_Ct(window, shape):
    B, G, F, H, W = shape
    edges = O.window_edges(window)
    Ne = len(edges)
    shp, win = L.make_shape(*shape), L.make_window(edges)
    x = rnd(B, G, F, H, W).requires_grad_(True)
    w = torch.softmax(rnd(B, G, Ne, H, W), dim=2).requires_grad_(True)
    g5, g6 = rnd(B, G, F, H, W), rnd(B, G, F, Ne, H, W)
    # L
    ref = O.op_L(x, w, edges)
    gx_r, gw_r = torch.autograd.grad(ref, [x, w], g5)
    out, gx, gw = torch.empty_like(x), torch.empty_like(x), torch.empty_like(w)
    E.call("glrgtv_op_L_fwd", shp, win, x.detach(), w.detach(), out, None)
    E.call("glrgtv_op_L_bwd", shp, win, x.detach(), w.detach(), g5, gx, gw, None)
    assert rel(out, ref) < 1e-6 and rel(gx, gx_r) < 1e-6 and rel(gw, gw_r) < 1e-6
    # C
    ref = O.op_C_core(x, w, edges)
    gx_r, gw_r = torch.autograd.grad(ref, [x, w], g6)
    z = torch.empty_like(g6)
    E.call("glrgtv_op_C_fwd", shp, win, x.detach(), w.detach(), z, None)
    E.call("glrgtv_op_C_bwd", shp, win, x.detach(), w.detach(), g6, gx, gw, None)
    assert rel(z, ref) < 1e-6 and rel(gx, gx_r) < 1e-6 and rel(gw, gw_r) < 1e-6
    # Ct
    zz = rnd(B, G, F, Ne, H, W).requires_grad_(True)
    ref = O.op_Ct_core(zz, w, edges)
    gz_r, gw_r = torch.autograd.grad(ref, [zz, w], g5)
    o, gz = torch.empty_like(x), torch.empty_like(g6)
    E.call("glrgtv_op_Ct_fwd", shp, win, zz.detach(), w.detach(), o, None)
    E.call("glrgtv_op_Ct_bwd", shp, win, zz.detach(), w.detach(), g5, gz, gw, None)
    assert rel(o, ref) < 1e-6 and rel(gz, gz_r) < 1e-6 and rel(gw, gw_r) < 1e-6


@pytest.mark.parametrize("shape", SHAPES[:3])
def test_soft_threshold_and_pooling(shape):
    B, G, F, H, W = shape
    Ne = 4
    shp = L.make_shape(*shape)
    t = rnd(B, G, F, Ne, H, W).requires_grad_(True)
    thr = (0.2 + torch.rand(G)).requires_grad_(True)
    g = rnd(B, G, F, Ne, H, W)
    ref = O.soft_threshold(t, thr)
    gt_r, gthr_r = torch.autograd.grad(ref, [t, thr], g)
    out, gt, gthr = torch.empty_like(g), torch.empty_like(g), torch.zeros(G)
    E.call("glrgtv_soft_threshold_fwd", shp, Ne, t.detach(), thr.detach(), out, None)
    E.call("glrgtv_soft_threshold_bwd", shp, Ne, t.detach(), thr.detach(), g, gt, gthr, None)
    assert torch.equal(out, ref.detach()) and torch.equal(gt, gt_r) and rel(gthr, gthr_r) < 1e-5
    if H % 2 == 0 and W % 2 == 0:
        x = rnd(B, G, F, H, W)
        c = torch.empty(B, G, F, H // 2, W // 2)
        E.call("glrgtv_pool2_fwd", shp, x, c, None)
        assert rel(c, O.pool2(x)) < 1e-6
        f = torch.empty_like(x)
        E.call("glrgtv_unpool2_fwd", shp, c, f, None)
        assert rel(f, O.unpool2(c)) < 1e-7
