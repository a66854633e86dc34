"""BASELINE config 5: the generalised unrolled block (older family's MixtureGTV solver) on one 512x512 image, C = 48 features
in G = 8 graphs, sweeping the unrolling depth {4, 8, 16, 32} and the stencil {3x3 cross, 3x3 full, 5x5 small, 5x5 full, 7x7 full}.
Runs on the per-operator kernels (`MixtureGTV.unrolled_solve(schedule=(n/2, n/2))`: one threshold step in the middle);
edge-weight extraction is inside the timed region.  Prints one JSON line per point (forward, no_grad): ms and Mpix/s.

    python tools/bench_config5.py [--res 512 --reps 3]

No import of oracle/: the window masks are restated here."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagerestoration_development_unrolling_b200 import model_GLR_GTV_deep_v7 as M  # noqa: E402


def full(n):
    w = np.ones((n, n), dtype=np.int64)
    w[n // 2, n // 2] = 0
    return w


WINDOWS = {"cross3": np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]]), "full3": full(3),
           "small5": np.array([[0, 0, 1, 0, 0], [0, 1, 1, 1, 0], [1, 1, 0, 1, 1], [0, 1, 1, 1, 0], [0, 0, 1, 0, 0]]), "full5": full(5), "full7": full(7)}


def flag(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    R, reps, G, F = flag("--res", 512), flag("--reps", 3), 8, 6
    dev = torch.device("cuda")
    z = lambda v: torch.tensor([[v], [0.0], [0.0], [0.0]])   # noqa: E731
    gen = torch.Generator().manual_seed(0)
    feat = torch.randn(1, G, F, R, R, generator=gen).to(dev)
    y = torch.randn(1, 1, 3, R, R, generator=gen).to(dev)
    for wname, mask in WINDOWS.items():
        for n_it in (4, 8, 16, 32):
            m = M.MixtureGTV(nchannels_in=3, n_graphs=G, n_node_fts=F, n_cnn_fts=8, connection_window=mask, n_cgd_iters=n_it, alpha_init=0.5,
                             beta_init=0.1, muy_init=z(0.03), ro_init=z(0.03), gamma_init=z(0.05), device=torch.device("cpu")).to(dev)
            schedule = (n_it // 2, n_it // 2)

            def run():
                with torch.no_grad():
                    wT, wL = m.GTVmodule00.extract_edge_weights(feat), m.GLRmodule00.extract_edge_weights(feat)
                    return m.unrolled_solve(y, wT, wL, schedule)

            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                out = run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(json.dumps({"window": wname, "edges": int(mask.sum()), "iterations": n_it, "ms": ms, "Mpix_per_s": R * R / ms / 1e3,
                              "finite": bool(torch.isfinite(out).all())}), flush=True)


if __name__ == "__main__":
    main()
