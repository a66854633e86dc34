"""Older model family (LIB/model_GLR_GTV_deep_v7.py): oracle pinned to the reference on CPU; drop-in module on GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import glr_gtv_oracle as O
from tests.util import rel

WINDOW = np.array([[0, 0, 1, 0, 0], [0, 1, 1, 1, 0], [1, 1, 0, 1, 1], [0, 1, 1, 1, 0], [0, 0, 1, 0, 0]])


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "v7_mixturegtv_g4.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_oracle_solver_matches_reference_fp64(golden_dir):
    g = _golden(golden_dir)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    out = O.mixture_gtv_solver(sd, g["x"], g["feats"], g["dc"], g["score"], window="small5")
    assert rel(out, g["out"]) < 1e-13


def test_state_dict_layout_matches_reference(golden_dir):
    from imagerestoration_development_unrolling_b200 import model_GLR_GTV_deep_v7 as M
    g = _golden(golden_dir)
    ref_keys = [k[3:] for k in g if k.startswith("sd.")]
    z = lambda v: torch.tensor([[v], [0.0], [0.0], [0.0]])
    m = M.MixtureGTV(nchannels_in=3, n_graphs=4, n_node_fts=3, n_cnn_fts=8, connection_window=WINDOW, n_cgd_iters=4,
                     alpha_init=0.5, beta_init=0.1, muy_init=z(0.1), ro_init=z(0.1), gamma_init=z(0.001), device=torch.device("cpu"))
    assert list(m.state_dict().keys()) == ref_keys
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(g["sd." + k].shape), k
    assert [tuple(d) for d in m.GTVmodule00.edge_delta.tolist()] == O.window_edges("small5")
    # default init of the shipped v7 model (V7:1047-1060)
    full = M.MultiScaleSequenceDenoiser(torch.device("cpu"))
    blk = full.mixtureGLR_block03
    assert blk.n_graphs == 24 and blk.GTVmodule00.n_edges == 12
    assert torch.allclose(blk.ro00, torch.full((24,), 0.1)) and torch.allclose(blk.gamma00, torch.log(torch.full((24,), 0.001)))
    assert torch.allclose(full.skip_connect_weight03, torch.tensor([0.1, 0.9]))


@pytest.mark.gpu
def test_mixture_gtv_module_matches_reference_golden(golden_dir):
    """whole MixtureGTV (CNN + graph solver) with the reference's weights, output and gradients (tolerance 1e-4)"""
    from imagerestoration_development_unrolling_b200 import model_GLR_GTV_deep_v7 as M
    g = _golden(golden_dir)
    z = lambda v: torch.tensor([[v], [0.0], [0.0], [0.0]])
    m = M.MixtureGTV(nchannels_in=3, n_graphs=4, n_node_fts=3, n_cnn_fts=8, connection_window=WINDOW, n_cgd_iters=4,
                     alpha_init=0.5, beta_init=0.1, muy_init=z(0.1), ro_init=z(0.1), gamma_init=z(0.001), device=torch.device("cpu"))
    m.load_state_dict({k[3:]: v.float() for k, v in g.items() if k.startswith("sd.")}, strict=True)
    m = m.cuda()
    x = g["x"].float().cuda().requires_grad_(True)
    out = m(x)
    names = [k for k, _ in m.named_parameters()]
    grads = torch.autograd.grad(out, [x] + [p for _, p in m.named_parameters()], g["gout"].float().cuda(), allow_unused=True)
    assert rel(out, g["out"]) < 1e-4
    assert rel(grads[0], g["gx"]) < 1e-4
    for k, gr in zip(names, grads[1:]):
        if "grad." + k in g and float(g["grad." + k].abs().max()) > 0:
            assert rel(gr, g["grad." + k]) < 3e-4, (k, rel(gr, g["grad." + k]))


@pytest.mark.gpu
@pytest.mark.parametrize("window", ["full3", "small5", "full5"])
def test_family_a_operators_against_oracle(window):
    """GLRFast / GTVFast of the older family: any window, reflect S, scalar stats, broadcast signal"""
    from imagerestoration_development_unrolling_b200 import model_GLR_GTV_deep_v7 as M
    mask = np.array(O.WINDOWS[window])
    edges = O.window_edges(window)
    G, Fn, B, H, W = 5, 3, 2, 9, 11
    dev = torch.device("cuda")
    glr, gtv = M.GLRFast(3, Fn, G, mask, dev, 1.0), M.GTVFast(3, Fn, G, mask, dev, 1.0)
    gen = torch.Generator().manual_seed(len(edges))
    with torch.no_grad():
        for mod in (glr, gtv):
            for p in mod.parameters():
                p.add_((0.3 * torch.randn(p.shape, generator=gen)).to(dev))
    feat = torch.randn(B, G, Fn, H, W, generator=gen)
    y = torch.randn(B, 1, 3, H, W, generator=gen)
    st = lambda mod: tuple(p.detach().double().cpu() for p in mod._stats())
    w_ref = O.edge_weights(feat.double(), gtv.multiM.detach().double().cpu(), edges)
    w, deg = gtv.extract_edge_weights(feat.to(dev))
    assert rel(w, w_ref) < 1e-5 and float((deg - 1).abs().max()) < 1e-5
    yG = y.double().expand(B, G, 3, H, W)
    assert rel(gtv.op_C(y.to(dev), w, deg), O.op_C(yG, w_ref, st(gtv), edges, "reflect")) < 1e-5
    assert rel(gtv(y.to(dev).expand(B, G, 3, H, W).contiguous(), w, deg), O.gtv_forward(yG, w_ref, st(gtv), edges, "reflect")) < 1e-5
    assert rel(glr(y.to(dev).expand(B, G, 3, H, W).contiguous(), w, deg), O.glr_forward(yG, w_ref, st(glr), edges, "reflect")) < 1e-5


# depths 4 / 6 / 8 / 16 / 32 (BASELINE config 5 sweeps the iteration count to 32); the deep schedules also vary the number of
# iterations between the ADMM passes
SWEEP = [("cross3", (2, 2)), ("full3", (2, 2, 2, 2)), ("small5", (2, 4)), ("full5", (2, 2, 2)), ("full7", (2, 2)),
         ("cross3", (2,) * 8), ("small5", (4, 4, 4, 4)), ("full5", (2, 6, 8)), ("cross3", (2,) * 16), ("full3", (8, 8, 8, 8)),
         ("full7", (4, 12, 16))]


@pytest.mark.gpu
@pytest.mark.parametrize("window,schedule", SWEEP)
def test_iteration_and_window_sweep_against_generalised_oracle(window, schedule):
    """BASELINE config 5 (iteration count / stencil sweep): the reference hard-codes both, so parity is
    against the loop-generalised oracle restatement, which is pinned to the reference at schedule (2,2) / 5x5-small by
    test_oracle_solver_matches_reference_fp64."""
    from imagerestoration_development_unrolling_b200 import model_GLR_GTV_deep_v7 as M
    mask, edges = np.array(O.WINDOWS[window]), O.window_edges(window)
    G, Fn, B, H, W, n_it = 3, 3, 1, 20, 24, sum(schedule)
    dev = torch.device("cuda")
    z = lambda v: torch.tensor([[v], [0.0], [0.0], [0.0]])
    m = M.MixtureGTV(nchannels_in=3, n_graphs=G, n_node_fts=Fn, n_cnn_fts=8, connection_window=mask, n_cgd_iters=n_it,
                     alpha_init=0.5, beta_init=0.1, muy_init=z(0.03), ro_init=z(0.03), gamma_init=z(0.05), device=torch.device("cpu"))
    gen = torch.Generator().manual_seed(n_it + len(edges))
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.startswith(("GTVmodule00", "GLRmodule00", "alphaCGD", "betaCGD")):
                p.add_(0.05 * torch.randn(p.shape, generator=gen))
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    m = m.to(dev)
    feat = torch.randn(B, G, Fn, H, W, generator=gen)
    y = torch.randn(B, 1, 3, H, W, generator=gen)
    wT_ref = O.edge_weights(feat.double(), sd["GTVmodule00.multiM"], edges)
    wL_ref = O.edge_weights(feat.double(), sd["GLRmodule00.multiM"], edges)
    ref = O.unrolled_admm_solve(sd, y.double().expand(B, G, 3, H, W), wT_ref, wL_ref, edges, schedule)
    wT, wL = m.GTVmodule00.extract_edge_weights(feat.to(dev)), m.GLRmodule00.extract_edge_weights(feat.to(dev))
    out = m.unrolled_solve(y.to(dev), wT, wL, schedule)
    assert rel(out, ref) < 1e-4, rel(out, ref)


# ---- the loop-generalised solver pinned at the reference's other fixed points: v1 = schedule (2, 4), full 3x3 / 5x5
#      windows, no stats convolutions (= identity stats: p01 = 1, p02a = p02b = p03 = 0)
def _v1_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    one, zero = torch.ones(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64)
    for mod in ("GTVmodule00.", "GLRmodule00."):
        sd.update({mod + "stats_kernel_p01": one, mod + "stats_kernel_p02a": zero, mod + "stats_kernel_p02b": zero, mod + "stats_kernel_p03": zero})
    return g, sd


def _v1_solve(g, sd):
    edges = O.window_edges(g["window"].tolist())
    B, _, H, W = g["x"].shape
    G, F = sd["GTVmodule00.multiM"].shape
    feats = g["feats"].reshape(B, G, F, H, W)
    wT, wL = O.edge_weights(feats, sd["GTVmodule00.multiM"], edges), O.edge_weights(feats, sd["GLRmodule00.multiM"], edges)
    out = O.unrolled_admm_solve(sd, g["x"][:, None].expand(B, G, 3, H, W), wT, wL, edges, schedule=(2, 4))
    return (out * g["score"][:, :, None]).sum(dim=1)


@pytest.mark.parametrize("name,n_edges", [("v1_mixturegtv_full3", 8), ("v1_mixturegtv_full5", 24)])
def test_oracle_generalised_solver_matches_reference_v1(golden_dir, name, n_edges):
    """outputs and the gradients of every graph parameter (the CNN outputs are inputs here, so these are total derivatives)"""
    g, sd = _v1_case(golden_dir, name)
    assert len(O.window_edges(g["window"].tolist())) == n_edges
    graph = ("ro00", "muys00", "gamma00", "alphaCGD", "betaCGD", "GTVmodule00.multiM", "GLRmodule00.multiM")
    for k in graph:
        sd[k] = sd[k].clone().requires_grad_(True)
    out = _v1_solve(g, sd)
    assert rel(out, g["out"]) < 1e-13
    grads = torch.autograd.grad(out, [sd[k] for k in graph], g["gout"])
    for k, gr in zip(graph, grads):
        ref = g["grad." + k]
        if float(ref.abs().max()) == 0.0:
            assert float(gr.abs().max()) == 0.0, k            # betaCGD rows 0 and 2: the first iteration of a solve has no momentum
        else:
            assert rel(gr, ref) < 1e-11, k


def test_oracle_generalised_solver_matches_reference_glr_only(golden_dir):
    """the GLR-only ablation (v13_no_orders_noGTV.GLR): schedule (3,), ro = 0, identity stats, exp(muys00), 3x3-cross window"""
    z = np.load(os.path.join(golden_dir, "glr_only_g3_f4.npz"))
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    G, F = g["sd.GLRmodule00.multiM"].shape
    B, C, H, W = g["x"].shape
    graph = ("muys00", "alphaCGD", "betaCGD", "GLRmodule00.multiM")
    p = {k: g["sd." + k].clone().requires_grad_(True) for k in graph}
    one, zero = torch.ones(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64)
    sd = {"alphaCGD": p["alphaCGD"], "betaCGD": p["betaCGD"], "muys00": p["muys00"].exp(), "ro00": torch.zeros(G, dtype=torch.float64),
          "gamma00": torch.zeros(G, dtype=torch.float64), "GLRmodule00.multiM": p["GLRmodule00.multiM"],
          "GTVmodule00.multiM": torch.ones(G, F, dtype=torch.float64)}
    for mod in ("GTVmodule00.", "GLRmodule00."):
        sd.update({mod + "stats_kernel_p01": one, mod + "stats_kernel_p02a": zero, mod + "stats_kernel_p02b": zero, mod + "stats_kernel_p03": zero})
    edges = O.window_edges("cross3")
    wL = O.edge_weights(g["feats"].reshape(B, G, F, H, W), sd["GLRmodule00.multiM"], edges)
    out = O.unrolled_admm_solve(sd, g["x"].reshape(B, G, F, H, W), torch.zeros_like(wL), wL, edges, schedule=(3,)).reshape(B, C, H, W)
    assert rel(out, g["out"]) < 1e-13
    for k, gr in zip(graph, torch.autograd.grad(out, [p[k] for k in graph], g["gout"])):
        ref = g["grad." + k]
        mask = ref != 0
        assert torch.equal(gr != 0, mask) and rel(gr[mask], ref[mask]) < 1e-11, k


# ---- the v1 three-block chain
def test_v1_denoiser_state_dict_layout(golden_dir):
    """same keys and parameter count as the reference's v1 MultiScaleSequenceDenoiser (recorded by make_golden_v1_denoiser.py)"""
    from imagerestoration_development_unrolling_b200 import model_GLR_GTV_deep_v1 as M1
    m = M1.MultiScaleSequenceDenoiser(torch.device("cpu"))
    z = np.load(os.path.join(golden_dir, "v1_denoiser.npz"))
    assert sorted(m.state_dict().keys()) == [str(k) for k in z["keys"]]
    assert sum(p.numel() for p in m.parameters()) == int(z["n_params"]) == 8282532
    assert [m.mixtureGLR_block01.GTVmodule00.n_edges, m.mixtureGLR_block03.GTVmodule00.n_edges] == [8, 24]


@pytest.mark.gpu
def test_v1_three_block_denoiser_matches_reference(golden_dir):
    """forward and backward of the whole v1 chain (three MixtureGTV blocks, 8 / 8 / 24 edges, schedule 2 + 4, SharpeningBlocks
    between) against the reference run in fp64 on the same deterministic parameters (tests/golden/v1_fill.py)"""
    import sys
    sys.path.insert(0, golden_dir)
    from v1_fill import fill_
    from imagerestoration_development_unrolling_b200 import model_GLR_GTV_deep_v1 as M1
    z = np.load(os.path.join(golden_dir, "v1_denoiser.npz"))
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        m = fill_(M1.MultiScaleSequenceDenoiser(torch.device("cpu"))).cuda()
        x = torch.from_numpy(z["x"]).float().cuda().requires_grad_(True)
        out = m(x)
        keys = [k[5:] for k in z.files if k.startswith("grad.")]
        params = dict(m.named_parameters())
        grads = torch.autograd.grad(out, [x] + [params[k] for k in keys], torch.from_numpy(z["gout"]).float().cuda())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert rel(out, torch.from_numpy(z["out"])) < 1e-4, rel(out, torch.from_numpy(z["out"]))
    assert rel(grads[0], torch.from_numpy(z["gx"])) < 1e-3, rel(grads[0], torch.from_numpy(z["gx"]))
    for k, g in zip(keys, grads[1:]):
        ref = torch.from_numpy(z["grad." + k])
        assert rel(g, ref) < 2e-3 or float((g.cpu().double() - ref).abs().max()) < 1e-6 * max(1.0, float(ref.abs().max())), (k, rel(g, ref))
