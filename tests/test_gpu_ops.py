"""Per-operator CUDA kernels (through ops.py -> C ABI) against the oracle and its autograd, on the GPU."""
import pytest
import torch

from oracle import glr_gtv_oracle as O

pytestmark = pytest.mark.gpu
WINDOWS = ["cross3", "full3", "small5", "full5"]
SHAPES = [(2, 2, 3, 5, 7), (1, 3, 6, 2, 2), (1, 1, 4, 1, 6), (2, 4, 6, 33, 47), (1, 2, 12, 64, 96)]


@pytest.fixture(scope="module")
def ops():
    from imagerestoration_development_unrolling_b200 import ops as o
    return o


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rnd(*s, seed=0):
    return torch.randn(*s, generator=torch.Generator().manual_seed(seed + sum(s)), dtype=torch.float64)


def leaf(t):
    return t.float().cuda().requires_grad_(True)


@pytest.mark.parametrize("window", WINDOWS)
@pytest.mark.parametrize("shape", SHAPES)
def test_edge_weights(ops, window, shape):
    B, G, F, H, W = shape
    edges = O.window_edges(window)
    feat, M = rnd(B, G, F, H, W).requires_grad_(True), (1 + 0.5 * rnd(G, F)).requires_grad_(True)
    gw = rnd(B, G, len(edges), H, W, seed=1)
    ref = O.edge_weights(feat, M, edges)
    gf_r, gM_r = torch.autograd.grad(ref, [feat, M], gw)
    f, m = leaf(feat), leaf(M)
    w = ops.edge_weights(f, m, ops.flat_edges(edges))
    gf, gM = torch.autograd.grad(w, [f, m], gw.float().cuda())
    assert rel(w, ref) < 2e-6 and rel(gf, gf_r) < 2e-5 and rel(gM, gM_r) < 2e-5


@pytest.mark.parametrize("pad", [0, 1])
@pytest.mark.parametrize("n_is_C", [True, False])
@pytest.mark.parametrize("shape", SHAPES)
def test_stats_conv(ops, pad, n_is_C, shape):
    B, G, F, H, W = shape
    if pad == 1 and (H < 2 or W < 2):
        pytest.skip("reflect needs >= 2 pixels")
    n = G * F if n_is_C else 1
    ps = [(torch.full((n, 1, 1, 1), v, dtype=torch.float64) + 0.2 * rnd(n, 1, 1, 1, seed=i)).requires_grad_(True)
          for i, v in enumerate((1.0, 0.5, 0.5, 0.5))]
    x, g = rnd(B, G, F, H, W).requires_grad_(True), rnd(B, G, F, H, W, seed=3)
    for op, ref_fn in ((ops.stats_conv, lambda: O.stats_conv(x, ps, "clamp" if pad == 0 else "reflect")),
                       (ops.stats_conv_t, lambda: O.stats_conv_transpose(x, ps))):
        ref = ref_fn()
        gr = torch.autograd.grad(ref, [x] + ps, g)
        xs, pc = leaf(x), [leaf(p) for p in ps]
        out = op(xs, *pc, pad)
        gg = torch.autograd.grad(out, [xs] + pc, g.float().cuda())
        assert rel(out, ref) < 1e-6
        # a scalar stats parameter's gradient is ONE sum over every pixel of the tensor, accumulated with fp32 atomics in an order
        # that changes from run to run (and cancels: 3.06e-5 was seen once on shape1); the parity bar for gradients is 1e-4
        for a, b in zip(gg, gr):
            assert rel(a, b) < (3e-5 if n_is_C else 1e-4)


@pytest.mark.parametrize("window", WINDOWS)
@pytest.mark.parametrize("shape", SHAPES)
def test_L_C_Ct(ops, window, shape):
    B, G, F, H, W = shape
    edges = O.window_edges(window)
    E, fe = len(edges), ops.flat_edges(edges)
    x = rnd(B, G, F, H, W).requires_grad_(True)
    w = torch.softmax(rnd(B, G, E, H, W, seed=1), dim=2).detach().requires_grad_(True)
    z = rnd(B, G, F, E, H, W, seed=2).requires_grad_(True)
    g5, g6 = rnd(B, G, F, H, W, seed=3), rnd(B, G, F, E, H, W, seed=4)
    for op, ref, a, g in ((ops.op_L, O.op_L(x, w, edges), x, g5), (ops.op_C, O.op_C_core(x, w, edges), x, g6),
                          (ops.op_Ct, O.op_Ct_core(z, w, edges), z, g5)):
        gr = torch.autograd.grad(ref, [a, w], g)
        ac, wc = leaf(a), leaf(w)
        out = op(ac, wc, fe)
        gg = torch.autograd.grad(out, [ac, wc], g.float().cuda())
        assert rel(out, ref) < 1e-6 and rel(gg[0], gr[0]) < 1e-6 and rel(gg[1], gr[1]) < 2e-6


@pytest.mark.parametrize("shape", SHAPES)
def test_soft_pool_normalize_gather(ops, shape):
    B, G, F, H, W = shape
    t, thr = rnd(B, G, F, 4, H, W).requires_grad_(True), (0.2 + torch.rand(G, dtype=torch.float64)).requires_grad_(True)
    g = rnd(B, G, F, 4, H, W, seed=5)
    ref = O.soft_threshold(t, thr)
    gr = torch.autograd.grad(ref, [t, thr], g)
    tc, hc = leaf(t), leaf(thr)
    out = ops.soft_threshold(tc, hc)
    gg = torch.autograd.grad(out, [tc, hc], g.float().cuda())
    assert rel(out, ref) < 1e-6 and rel(gg[0], gr[0]) < 1e-6 and rel(gg[1], gr[1]) < 1e-4
    # normalise
    feat, M = rnd(B, G, F, H, W).requires_grad_(True), (1 + 0.5 * rnd(G, F)).requires_grad_(True)
    ref = O.normalize_transform(feat, M)
    gr = torch.autograd.grad(ref, [feat, M], rnd(B, G, F, H, W, seed=6))
    fc, mc = leaf(feat), leaf(M)
    out = ops.normalize_transform(fc, mc)
    gg = torch.autograd.grad(out, [fc, mc], rnd(B, G, F, H, W, seed=6).float().cuda())
    assert rel(out, ref) < 1e-6 and rel(gg[0], gr[0]) < 1e-5 and rel(gg[1], gr[1]) < 1e-5
    # gather
    edges = O.window_edges("small5")
    x = rnd(B, G, F, H, W).requires_grad_(True)
    ref = torch.stack([O.shift_clamp(x, dh, dw) for dh, dw in edges], dim=3)
    gg6 = rnd(B, G, F, len(edges), H, W, seed=7)
    gr = torch.autograd.grad(ref, x, gg6)[0]
    xc = leaf(x)
    out = ops.gather_neighbors(xc, ops.flat_edges(edges))
    assert torch.equal(out.cpu().double(), ref.detach().float().double())
    assert rel(torch.autograd.grad(out, xc, gg6.float().cuda())[0], gr) < 1e-6
    if H % 2 == 0 and W % 2 == 0:
        xc = leaf(x)
        c = ops.pool2(xc)
        assert rel(c, O.pool2(x)) < 1e-6
        assert rel(torch.autograd.grad(c, xc, torch.ones_like(c))[0], 0.25 * torch.ones_like(x)) < 1e-7
        assert rel(ops.unpool2(c), O.unpool2(O.pool2(x))) < 1e-6


def test_cpu_tensors_are_refused(ops):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.pool2(torch.zeros(1, 1, 1, 2, 2))


PROJ_SHAPES = [(2, 96, 48, 64 * 64), (1, 48, 192, 32 * 32), (3, 384, 1536, 16 * 16), (2, 768, 384, 256), (2, 192, 96, 1024),
               (3, 1536, 384, 64), (2, 24, 12, 36), (1, 8, 32, 20), (2, 96, 48, 336 * 496 // 16)]


@pytest.mark.timeout(120, method="thread")
@pytest.mark.parametrize("pipeline", [0, 1], ids=["in_place_stages", "landing_ring"])
@pytest.mark.parametrize("shape", PROJ_SHAPES)
def test_projection_gemm_3xtf32(shape, pipeline):
    """glrgtv_proj_gemm / glrgtv_proj_wgrad (tcgen05 kind::tf32, three-pass split, csrc/proj_tc.cu) against fp64 GEMMs:
    fp32-level accuracy for the forward, the input gradient and the weight gradient; ragged tiles (pixels not a multiple of
    128, channels not a multiple of 16 / 32) are zero-padded by the TMA unit"""
    from imagerestoration_development_unrolling_b200 import ops, _lib as L
    assert L.load().glrgtv_set_proj_pipeline(pipeline) == 0       # both pipelines of csrc/proj_tc.cu (0 is the default)
    try:
        _check_projection(ops, shape)
    finally:
        L.load().glrgtv_set_proj_pipeline(0)


def _check_projection(ops, shape):
    B, M, K, N = shape
    gen = torch.Generator().manual_seed(M + K)
    w = torch.randn(M, K, generator=gen).cuda().requires_grad_(True)
    x = torch.randn(B, K, N, generator=gen).cuda().requires_grad_(True)
    gy = torch.randn(B, M, N, generator=gen).cuda()
    assert ops.proj_supported(M, K, N)
    y = ops.proj_gemm(w, x, False)
    gw, gx = torch.autograd.grad(y, [w, x], gy)
    w64, x64, gy64 = w.detach().double(), x.detach().double(), gy.double()
    # three-pass TF32: products are exact to ~2^-21; the tensor core's fp32 accumulator truncates, which adds ~7e-9 per unit of K
    tol = 2e-6 + 1e-8 * max(M, K)
    assert rel(y, torch.einsum("mk,bkn->bmn", w64, x64)) < tol
    assert rel(gx, torch.einsum("mk,bmn->bkn", w64, gy64)) < tol
    assert rel(gw, torch.einsum("bmn,bkn->mk", gy64, x64)) < 5e-6
    # the transposed form on its own (W^T X), with its own gradients
    x2 = torch.randn(B, M, N, generator=gen).cuda().requires_grad_(True)
    g2 = torch.randn(B, K, N, generator=gen).cuda()
    y2 = ops.proj_gemm(w, x2, True)
    gw2, gx2 = torch.autograd.grad(y2, [w, x2], g2)
    assert rel(y2, torch.einsum("mk,bmn->bkn", w64, x2.detach().double())) < tol
    assert rel(gx2, torch.einsum("mk,bkn->bmn", w64, g2.double())) < tol
    assert rel(gw2, torch.einsum("bmn,bkn->mk", x2.detach().double(), g2.double())) < 5e-6


@pytest.mark.parametrize("shape", [(2, 6, 8, 16), (1, 48, 64, 256), (3, 5, 2, 8)])
def test_space_to_depth_matches_pixel_unshuffle(shape):
    from imagerestoration_development_unrolling_b200 import ops
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)).cuda().requires_grad_(True)
    y = ops.space_to_depth(x, False)
    ref = torch.nn.functional.pixel_unshuffle(x, 2)
    assert torch.equal(y, ref)
    g = torch.randn_like(ref)
    assert torch.equal(torch.autograd.grad(y, x, g)[0], torch.nn.functional.pixel_shuffle(g, 2))
    assert torch.equal(ops.space_to_depth(y.detach(), True), x.detach())
