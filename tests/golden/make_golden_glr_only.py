"""Golden vector pinning the oracle's loop-generalised solver at the reference's GLR-only ablation
(model_GLR_GTV_deep_v13_no_orders_noGTV.GLR: one solve of 3 momentum iterations, no GTV term, no stats convolutions,
3x3-cross window, exp(muys00)).  Produced by the REFERENCE (CPU, fp64), unmodified.
Run in the build container only:   python tests/golden/make_golden_glr_only.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/exploration/model_multiscale_mixture_GLR/lib")
import model_GLR_GTV_deep_v13_no_orders_noGTV as ref  # noqa: E402  (the reference, unmodified)


def main():
    torch.manual_seed(21)
    m = ref.GLR(n_graphs=3, n_node_fts=4, alpha_init=0.5, beta_init=0.1, muy_init=torch.tensor([[0.3]]))
    gen = torch.Generator().manual_seed(22)
    graph = ("muys00", "alphaCGD", "betaCGD", "GLRmodule00.multiM")
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k in graph:
                p.add_(0.2 * torch.randn(p.shape, generator=gen))
    m.double()
    x = torch.randn(2, 12, 10, 14, generator=gen, dtype=torch.float64)
    gout = torch.randn(2, 12, 10, 14, generator=gen, dtype=torch.float64)
    cap = {}
    m.patchs_features_extraction00.register_forward_hook(lambda mod, i, o: cap.__setitem__("feats", o.detach()))
    out = m(x)
    params = dict(m.named_parameters())
    grads = torch.autograd.grad(out, [params[k] for k in graph], gout)
    rec = {"x": x.numpy(), "gout": gout.numpy(), "out": out.detach().numpy(), "feats": cap["feats"].numpy()}
    for k, gr in zip(graph, grads):
        rec["sd." + k] = params[k].detach().numpy()
        rec["grad." + k] = gr.numpy()
    np.savez_compressed(os.path.join(HERE, "glr_only_g3_f4.npz"), **rec)
    print("written; |out - x| / |x| =", float((out.detach() - x).norm() / x.norm()))


if __name__ == "__main__":
    main()
