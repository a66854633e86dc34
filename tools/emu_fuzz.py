"""Random-shape sweep of the streaming block kernels in the g++ emulation against the oracle (forward intermediates + output, backward input /
parameter / projection gradients), both loaders.  A development tool, slower than the unit tests:
    python tools/emu_fuzz.py [seed] [n_cases]          (also under GLRGTV_EMU_SCHED=... / GLRGTV_EMU_ASYNC=late, tools/emu_races.sh)
    python tools/emu_fuzz.py [seed] [n_cases] --plane  the shared-memory plane kernels instead (any even H, W >= 4)
Graphs with ONE feature are left out: their normalised feature is +-1, the edge weights do not depend on the projections, and the
reference gradient of the projection weights is rounding noise around zero (a relative comparison is meaningless there)."""
import random, sys, time, traceback
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import emu_harness as E
from tests import test_emu_stream as T
lib = E.emu_lib()
PLANE = "--plane" in sys.argv
if PLANE:
    sys.argv.remove("--plane")
    from tests import test_emu_block as TB
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
fails = []
t0 = time.time()
for it in range(n):
    G = rng.choice([1, 2, 3, 4]); F = rng.choice([2, 3, 4, 6]); B = rng.choice([1, 1, 2])
    H = 2 * rng.choice([1, 2, 3, 4, 5, 7, 9, 13, 17, 20, 33, 64])
    W = 8 * rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 17, 18, 31, 32, 34, 40])
    if H * W * G * F * B > 60000: H = max(2, 2 * (60000 // (W * G * F * B) // 2)); 
    case = (G * F, G, B, H, W)
    if PLANE:       # 2x2 images are left out: their 1x1 coarse plane makes the coarse GLR term rounding noise around zero
        case = (G * F, G, B, 2 * rng.choice([2, 3, 4, 5, 7, 9, 13, 16, 17, 20, 33]), 2 * rng.choice([2, 3, 5, 7, 9, 11, 15, 16, 17, 19, 31, 33, 35, 47, 65]))
        lib.glrgtv_set_block_path(1)
        for name, fn in (("fwd", TB.test_block_forward), ("bwd", TB.test_block_backward)):
            try:
                fn(case)
            except Exception as e:
                fails.append((case, 0, name, repr(e)[:300]))
                print("FAIL", case, name, repr(e)[:300], flush=True)
        continue
    for loader in (1, 2):
        lib.glrgtv_set_block_path(2); lib.glrgtv_set_stream_loader(loader)
        for name, fn in (("fwd", lambda: T.test_stream_block_forward(case)), ("bwd", lambda: T.test_stream_block_backward(case, 0))):
            try:
                fn()
            except Exception as e:
                fails.append((case, loader, name, repr(e)[:300]))
                print("FAIL", case, loader, name, repr(e)[:300], flush=True)
lib.glrgtv_set_block_path(0); lib.glrgtv_set_stream_loader(0)
print(f"{n} cases, {len(fails)} failures, {time.time()-t0:.0f}s")
sys.exit(1 if fails else 0)
