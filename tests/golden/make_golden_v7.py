"""Golden vectors for the older family, produced by the REFERENCE's model_GLR_GTV_deep_v7.MixtureGTV (CPU, fp64).
Run in the build container only:   python tests/golden/make_golden_v7.py
Stores the CNN outputs (features, DC term, mixture scores) next to the result so that the oracle's restatement of the
graph solver can be pinned without restating the out-of-scope CNN (V7:936-1016)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/exploration/model_multiscale_mixture_GLR/lib")
import model_GLR_GTV_deep_v7 as ref  # noqa: E402  (the reference, unmodified)

WINDOW = np.array([[0, 0, 1, 0, 0], [0, 1, 1, 1, 0], [1, 1, 0, 1, 1], [0, 1, 1, 1, 0], [0, 0, 1, 0, 0]])


def main():
    torch.manual_seed(3)
    dev = torch.device("cpu")
    z = lambda v: torch.tensor([[v], [0.0], [0.0], [0.0]])
    m = ref.MixtureGTV(nchannels_in=3, n_graphs=4, n_node_fts=3, n_cnn_fts=8, connection_window=WINDOW, n_cgd_iters=4,
                       alpha_init=0.5, beta_init=0.1, muy_init=z(0.1), ro_init=z(0.1), gamma_init=z(0.001), device=dev)
    gen = torch.Generator().manual_seed(4)
    with torch.no_grad():   # move the graph parameters off their init
        for name, p in m.named_parameters():
            if name.startswith("patchs_features_extraction") or name.startswith("dc_estimator") or name.startswith("combination"):
                continue
            if name in ("ro00", "muys00"):
                p.copy_(0.02 + 0.06 * torch.rand(p.shape, generator=gen))
            elif name == "gamma00":
                p.copy_(torch.log(0.02 + 0.2 * torch.rand(p.shape, generator=gen)))
            elif name.endswith("multiM"):
                p.copy_(1.0 + 0.5 * torch.randn(p.shape, generator=gen))
            else:
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
    m.double()
    for mod in m.modules():
        for n in ("stats_kernel01", "stats_kernel02a", "stats_kernel02b", "stats_kernel03"):
            if hasattr(mod, n):
                setattr(mod, n, getattr(mod, n).double())
    x = torch.rand(2, 3, 12, 16, generator=gen, dtype=torch.float64)
    gout = torch.randn(2, 3, 12, 16, generator=gen, dtype=torch.float64)
    cap = {}
    m.patchs_features_extraction.register_forward_hook(lambda mod, i, o: cap.__setitem__("feats", o[0].detach()))
    m.dc_estimator.register_forward_hook(lambda mod, i, o: cap.__setitem__("dc", o.detach()))
    m.combination_weight.register_forward_hook(lambda mod, i, o: cap.__setitem__("score", o.detach()))
    xx = x.clone().requires_grad_(True)
    out = m(xx)
    names = [k for k, _ in m.named_parameters()]
    grads = torch.autograd.grad(out, [xx] + [p for _, p in m.named_parameters()], gout, allow_unused=True)
    rec = {"x": x.numpy(), "gout": gout.numpy(), "out": out.detach().numpy(), "gx": grads[0].numpy(),
           "feats": cap["feats"].numpy(), "dc": cap["dc"].numpy(), "score": cap["score"].numpy()}
    for k, v in m.state_dict().items():
        rec["sd." + k] = v.numpy()
    for k, gr in zip(names, grads[1:]):
        if gr is not None and not k.startswith(("patchs_features_extraction", "dc_estimator", "combination")):
            rec["grad." + k] = gr.numpy()
    np.savez_compressed(os.path.join(HERE, "v7_mixturegtv_g4.npz"), **rec)
    print("written; |out - x| / |x| =", float((out.detach() - x).norm() / x.norm()))


if __name__ == "__main__":
    main()
