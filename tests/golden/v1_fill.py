"""Deterministic parameter fill shared by tests/golden/make_golden_v1_denoiser.py (applied to the REFERENCE model) and
tests/test_family_a.py (applied to this package's model): the v1 denoiser has 8.3 M parameters, too many to commit as a
fixture, so both sides regenerate them from the key names' order.  CPU torch.randn with a fixed seed per key."""
import torch


def fill_(module, seed=2204):
    """overwrite every parameter: convolution weights ~ N(0, 1/fan_in) (identity-ish depthwise norm scales kept near 1), graph
    parameters inside the ranges the solver is stable in (off their init)."""
    with torch.no_grad():
        for i, (k, p) in enumerate(sorted(module.state_dict().items())):
            g = torch.Generator().manual_seed(seed + i)
            r = torch.randn(p.shape, generator=g, dtype=torch.float32)
            u = torch.rand(p.shape, generator=g, dtype=torch.float32)
            leaf = k.rsplit(".", 1)[-1] if not k.endswith("weight") else k
            if k.endswith("weighted_transform.weight"):                       # CustomLayerNorm scale
                v = 1.0 + 0.1 * r
            elif k.endswith(".weight") and p.dim() == 4:
                v = r / float(p.shape[1] * p.shape[2] * p.shape[3]) ** 0.5
            elif leaf in ("ro00", "muys00"):
                v = 0.02 + 0.06 * u
            elif leaf == "gamma00":
                v = torch.log(0.02 + 0.2 * u)
            elif leaf == "multiM":
                v = 1.0 + 0.5 * r
            elif leaf in ("alphaCGD", "betaCGD"):
                v = (0.5 if leaf == "alphaCGD" else 0.1) + 0.1 * r
            elif "skip_connect_weight" in leaf:
                v = 0.5 + 0.2 * (u - 0.5)
            else:
                raise KeyError(f"v1_fill: no rule for {k} {tuple(p.shape)}")
            p.copy_(v.to(p.dtype))
    return module
