"""helpers shared by the tests (CPU and GPU)."""
import torch

from imagerestoration_development_unrolling_b200 import _lib as L
from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
from oracle import glr_gtv_oracle as O


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def random_block_state(dim, ngraphs, seed):
    """state-dict of a LocalLowpassFilteringBlock with every parameter moved off its init."""
    torch.manual_seed(seed)
    blk = M.LocalLowpassFilteringBlock(dim=dim, nsubnets=1, ngraphs=ngraphs)
    sd = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    return O.randomize_block_state(sd, seed + 1)


def block_structs(sd, prefix="local_filter.", skip=True, dev=None):
    """(BlockParams, keepalive) from a state dict of float32 contiguous tensors on one device."""
    t = {k: (v.to(dev) if dev is not None else v).contiguous().float() for k, v in sd.items()}
    p = L.BlockParams()
    for field, mod in (("gtv0", "GTVmodule00."), ("glr0", "GLRmodule00."), ("gtv1", "GTVmodule01."), ("glr1", "GLRmodule01.")):
        op = getattr(p, field)
        op.stats = L.make_stats(*[t[prefix + mod + n] for n in ("stats_kernel_p01", "stats_kernel_p02a", "stats_kernel_p02b", "stats_kernel_p03")])
        op.multiM = t[prefix + mod + "multiM"].data_ptr()
    for field, key in (("alpha", "alphaCGD"), ("beta", "betaCGD"), ("mu0", "muys00"), ("ro0", "ro00"), ("gamma0", "gamma00"),
                       ("mu1", "muys01"), ("ro1", "ro01"), ("gamma1", "gamma01")):
        setattr(p, field, t[prefix + key].data_ptr())
    p.skip = t["skip_weight"].data_ptr() if skip else None
    return p, t


def alloc_saved(B, G, F, H, W, dev="cpu"):
    sv = L.BlockSaved()
    keep = {}
    for n in ("wT0", "wL0"):
        keep[n] = torch.empty(B, G, 4, H, W, device=dev)
    for n in ("wT1", "wL1"):
        keep[n] = torch.empty(B, G, 4, H // 2, W // 2, device=dev)
    for n in ("bA", "x1", "bB", "r1", "x2"):
        keep[n] = torch.empty(B, G, F, H, W, device=dev)
    keep["cT0"] = torch.empty(B, G, 2, H, W, device=dev)
    keep["cT1"] = torch.empty(B, G, 2, H // 2, W // 2, device=dev)
    keep["vc"] = torch.empty(2, B, G, F, H // 2, W // 2, device=dev)
    for n, v in keep.items():
        setattr(sv, n, v.data_ptr())
    return sv, keep


def oracle_features(sd, x, prefix="local_filter."):
    """the two projections, done by the oracle's einsum convs (CPU)."""
    f0 = O._conv1x1(x, sd[prefix + "patchs_features_extraction00.0.weight"])
    f1 = O._conv1x1(O._conv2x2s2(x, sd[prefix + "patchs_features_extraction01.0.weight"]),
                    sd[prefix + "patchs_features_extraction01.1.weight"])
    return f0.contiguous(), f1.contiguous()


GRAD_FIELDS = (("gtv0_stats", "GTVmodule00."), ("glr0_stats", "GLRmodule00."), ("gtv1_stats", "GTVmodule01."),
               ("glr1_stats", "GLRmodule01."))
STATS_NAMES = ("stats_kernel_p01", "stats_kernel_p02a", "stats_kernel_p02b", "stats_kernel_p03")


def alloc_grads(G, F, dev="cpu", skip=True):
    """zeroed gradient buffers + the BlockGrads struct pointing at them."""
    C = G * F
    keep = {}
    for f, _ in GRAD_FIELDS:
        keep[f] = torch.zeros(4 * C, device=dev)
    for f in ("gtv0_M", "glr0_M", "gtv1_M", "glr1_M"):
        keep[f] = torch.zeros(G, F, device=dev)
    for f in ("alpha", "beta"):
        keep[f] = torch.zeros(3, G, device=dev)
    for f in ("mu0", "ro0", "gamma0", "mu1", "ro1", "gamma1"):
        keep[f] = torch.zeros(G, device=dev)
    keep["skip"] = torch.zeros(2, device=dev)
    gr = L.BlockGrads()
    for n, v in keep.items():
        setattr(gr, n, v.data_ptr() if (n != "skip" or skip) else None)
    return gr, keep


def grads_to_state_names(keep, G, F, prefix="local_filter."):
    """map the C-ABI gradient buffers onto the state-dict parameter names."""
    C = G * F
    out = {}
    for f, mod in GRAD_FIELDS:
        for i, n in enumerate(STATS_NAMES):
            out[prefix + mod + n] = keep[f][i * C:(i + 1) * C].reshape(C, 1, 1, 1)
        out[prefix + mod + "multiM"] = keep[f.replace("_stats", "_M")]
    for f, k in (("alpha", "alphaCGD"), ("beta", "betaCGD"), ("mu0", "muys00"), ("ro0", "ro00"), ("gamma0", "gamma00"),
                 ("mu1", "muys01"), ("ro1", "ro01"), ("gamma1", "gamma01")):
        out[prefix + k] = keep[f]
    out["skip_weight"] = keep["skip"]
    return out
