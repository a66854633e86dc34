"""Golden vectors pinning the loop-generalised solver of the oracle (oracle.unrolled_admm_solve) at the reference's OTHER fixed
points: model_GLR_GTV_deep_v1.MixtureGTV - schedule (2, 4) = 6 iterations, full 3x3 window (8 edges) and full 5x5
window (24 edges), no stats convolutions.  Produced by the REFERENCE (CPU, fp64), unmodified.
Run in the build container only:   python tests/golden/make_golden_v1.py
The CNN outputs (features, mixture scores) are stored next to the result so that the graph solver can be pinned without
restating the out-of-scope feature CNN (v1:602-676)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/exploration/model_multiscale_mixture_GLR/lib")
import model_GLR_GTV_deep_v1 as ref  # noqa: E402  (the reference, unmodified)


def full_window(n):
    w = np.ones((n, n), dtype=np.int64)
    w[n // 2, n // 2] = 0
    return w


def case(name, window, seed):
    torch.manual_seed(seed)
    dev = torch.device("cpu")
    z = lambda v: torch.tensor([[v], [0.0], [0.0], [0.0]])
    m = ref.MixtureGTV(nchannels_in=3, n_graphs=4, n_node_fts=3, connection_window=window, n_cgd_iters=6, alpha_init=0.5,
                       beta_init=0.1, muy_init=z(0.1), ro_init=z(0.1), gamma_init=z(0.001), device=dev)
    gen = torch.Generator().manual_seed(seed + 1)
    graph = ("ro00", "muys00", "gamma00", "alphaCGD", "betaCGD", "GTVmodule00.multiM", "GLRmodule00.multiM")
    with torch.no_grad():   # move the graph parameters off their init
        for k, p in m.named_parameters():
            if k not in graph:
                continue
            if k in ("ro00", "muys00"):
                p.copy_(0.02 + 0.06 * torch.rand(p.shape, generator=gen))
            elif k == "gamma00":
                p.copy_(torch.log(0.02 + 0.2 * torch.rand(p.shape, generator=gen)))
            elif k.endswith("multiM"):
                p.copy_(1.0 + 0.5 * torch.randn(p.shape, generator=gen))
            else:
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
    m.double()
    x = torch.rand(2, 3, 16, 24, generator=gen, dtype=torch.float64)
    gout = torch.randn(2, 3, 16, 24, generator=gen, dtype=torch.float64)
    cap = {}
    m.patchs_features_extraction.register_forward_hook(lambda mod, i, o: cap.__setitem__("feats", o[0].detach()))
    m.combination_weight.register_forward_hook(lambda mod, i, o: cap.__setitem__("score", o.detach()))
    out = m(x)
    params = dict(m.named_parameters())
    grads = torch.autograd.grad(out, [params[k] for k in graph], gout)
    rec = {"x": x.numpy(), "gout": gout.numpy(), "out": out.detach().numpy(), "feats": cap["feats"].numpy(), "score": cap["score"].numpy(),
           "window": window}
    for k, gr in zip(graph, grads):
        rec["sd." + k] = params[k].detach().numpy()
        rec["grad." + k] = gr.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, "written; |out - x| / |x| =", float((out.detach() - x).norm() / x.norm()))


if __name__ == "__main__":
    case("v1_mixturegtv_full3", full_window(3), 11)
    case("v1_mixturegtv_full5", full_window(5), 12)
