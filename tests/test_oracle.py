"""The oracle (oracle/glr_gtv_oracle.py) against golden vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import glr_gtv_oracle as O

BLOCK_CASES = ["block_f6_g2", "block_f12_g2", "block_f6_g4_tiny"]
OP_CASES = ["ops_f6_g2", "ops_f3_g3_small"]


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


@pytest.mark.parametrize("name", BLOCK_CASES)
def test_block_forward_backward_fp64(golden_dir, name):
    g = _load(golden_dir, name)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    out, gx, pg = O.lowpass_block_fwd_bwd(sd, g["x"], g["gout"])
    assert rel(out, g["out"]) < 1e-13
    assert rel(gx, g["gx"]) < 1e-12
    for k in sd:
        ref = g["grad." + k]
        if float(ref.abs().max()) == 0.0:
            assert float(pg[k].abs().max()) == 0.0, k     # betaCGD rows 0,1 are exact zeros
        else:
            assert rel(pg[k], ref) < 1e-11, (k, rel(pg[k], ref))


@pytest.mark.parametrize("name", BLOCK_CASES)
def test_block_forward_fp32_noise_floor(golden_dir, name):
    """oracle in fp32 is as close to the fp64 answer as the reference's own fp32 run is (x10 slack)."""
    g = _load(golden_dir, name)
    sd = {k[3:]: v.float() for k, v in g.items() if k.startswith("sd.")}
    out = O.lowpass_block_forward(sd, g["x"].float()).double()
    floor = rel(g["out32"], g["out"])
    assert rel(out, g["out"]) < max(10 * floor, 1e-6)


@pytest.mark.parametrize("name", OP_CASES)
def test_operators_fp64(golden_dir, name):
    g = _load(golden_dir, name)
    edges = O.window_edges("cross3")
    assert edges == [(-1, 0), (0, -1), (0, 1), (1, 0)]
    x, feat, z = g["x"], g["feat"], g["z"]
    for tag in ("glr", "gtv"):
        stats = tuple(g[f"{tag}.stats_kernel_p0{s}"] for s in ("1", "2a", "2b", "3"))
        assert rel(O.stats_conv(x, stats), g[f"{tag}.S"]) < 1e-14
        assert rel(O.stats_conv_transpose(x, stats), g[f"{tag}.St"]) < 1e-14
    sL = tuple(g[f"glr.stats_kernel_p0{s}"] for s in ("1", "2a", "2b", "3"))
    sT = tuple(g[f"gtv.stats_kernel_p0{s}"] for s in ("1", "2a", "2b", "3"))
    w_glr = O.edge_weights(feat, g["glr.multiM"], edges)
    w_gtv = O.edge_weights(feat * 0.7 + 0.1, g["gtv.multiM"], edges)
    assert rel(w_glr, g["w_glr"]) < 1e-14
    assert rel(w_gtv, g["w_gtv"]) < 1e-14
    assert rel(O.op_L(x, w_glr, edges), g["glr.L"]) < 1e-14
    assert rel(O.glr_forward(x, w_glr, sL, edges), g["glr.fwd"]) < 1e-14
    assert rel(O.op_C(x, w_gtv, sT, edges), g["gtv.C"]) < 1e-14
    assert rel(O.op_Ct(z, w_gtv, sT, edges), g["gtv.Ct"]) < 1e-14
    assert rel(O.gtv_forward(x, w_gtv, sT, edges), g["gtv.fwd"]) < 1e-14
    assert rel(O.soft_threshold(z, g["thr"]), g["soft"]) < 1e-15


def test_window_edge_order():
    # SURVEY 3.4 [probe]: 5x5-small window order
    assert O.window_edges("small5") == [(-2, 0), (-1, -1), (-1, 0), (-1, 1), (0, -2), (0, -1),
                                        (0, 1), (0, 2), (1, -1), (1, 0), (1, 1), (2, 0)]
    assert len(O.window_edges("full3")) == 8 and len(O.window_edges("full5")) == 24
