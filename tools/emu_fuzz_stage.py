"""Random sweep of glrgtv_block_fwd_stage in the g++ emulation: random block shapes, every stage run on a random partition of the rows
into even ranges, the ranges of a stage in random order (what spatially sharded callers do), NaN-filled buffers, against the oracle.
    python tools/emu_fuzz_stage.py [seed] [n_cases]"""
import random, sys, time, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import glr_gtv_oracle as O
from imagerestoration_development_unrolling_b200 import _lib as L
from tests import emu_harness as E
from tests.util import rel, random_block_state, block_structs, alloc_saved, oracle_features
lib = E.emu_lib(); rng = random.Random(int(sys.argv[1])); n = int(sys.argv[2]); fails = 0; t0 = time.time()
for it in range(n):
    G = rng.choice([1, 2, 3]); F = rng.choice([2, 3, 4]); B = rng.choice([1, 2]); dim = G * F
    H = 2 * rng.randint(2, 40); W = 8 * rng.choice([1, 2, 3, 5, 8, 9, 16, 17, 33, 34])
    if B * dim * H * W > 50000: H = max(4, 2 * (50000 // (B * dim * W) // 2))
    loader = rng.choice([1, 2]); lib.glrgtv_set_block_path(2); lib.glrgtv_set_stream_loader(loader)
    sd = random_block_state(dim, G, seed=it)
    x = torch.randn(B, dim, H, W, generator=torch.Generator().manual_seed(it))
    ref_out, inter = O.mixture_gtvglr_forward({k: v.double() for k, v in sd.items()}, x.double(), "local_filter.", return_intermediates=True)
    s = sd["skip_weight"].double(); ref_out = s[0] * x.double() + s[1] * ref_out
    f0, f1 = oracle_features(sd, x); p, keep = block_structs(sd); sv, saved = alloc_saved(B, G, F, H, W)
    for t in saved.values(): t.fill_(float("nan"))
    out = torch.full_like(x, float("nan")); shp = L.make_shape(B, G, F, H, W)
    E.call("glrgtv_block_fwd_stage", 0, shp, p, x, f0, f1, out, sv, 0, H, None)
    k = rng.randint(1, min(5, H // 2)); cuts = sorted(set([0, H] + [2 * rng.randint(1, H // 2 - 1) for _ in range(k - 1)])) if H > 4 else [0, H]
    order = list(zip(cuts[:-1], cuts[1:]))
    try:
        for stage in (1, 2, 3, 4):
            rng.shuffle(order)                                # ranges in any order within a stage
            for a, b in order: E.call("glrgtv_block_fwd_stage", stage, shp, p, x, None, None, out, sv, a, b, None)
        for nme in ("bA", "x1", "bB", "r1", "x2"): assert rel(saved[nme], inter[nme]) < 2e-5, nme
        assert rel(out, ref_out) < 1e-5
    except Exception as e:
        fails += 1; print("FAIL", (dim, G, B, H, W), loader, cuts, repr(e)[:200], flush=True)
lib.glrgtv_set_block_path(0); lib.glrgtv_set_stream_loader(0)
print(f"{n} cases, {fails} failures, {time.time()-t0:.0f}s")
sys.exit(1 if fails else 0)
