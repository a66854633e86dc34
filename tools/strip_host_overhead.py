"""How much of a rank's 4K-inference step is host time?  One GPU plays ONE rank of an N-rank run: the four filter blocks on row strips
of H/N rows through shard.sharded_filtering_staged (stage runners forced, no neighbours, so no exchange), timed three ways:
CUDA events around the call (GPU + host stalls), host time to ENQUEUE the call (no synchronisation), and the per-kernel CUDA-event
sum from glrgtv_profile_* (pure kernel time of the stage / weight kernels; projections not included).

    python tools/strip_host_overhead.py [--ranks 8] [--reps 10]"""
import argparse, ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagerestoration_development_unrolling_b200 import _lib as L, shard
from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M

ap = argparse.ArgumentParser()
ap.add_argument("--ranks", type=int, default=8)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda")
lib = L.load()
DIMS, NG, H0, W0 = [48, 96, 192, 384], [8, 16, 16, 32], 2160, 3840
torch.manual_seed(0)
blocks = [M.LocalLowpassFilteringBlock(d, 1, g).to(dev) for d, g in zip(DIMS, NG)]
strips = []
for s, d in enumerate(DIMS):
    r0, r1 = shard.strip_bounds(H0 >> s, a.ranks, align=2)[0]
    strips.append(torch.randn(1, d, r1 - r0 + 2 * shard.STAGE_HALO_ROWS, W0 >> s, device=dev))      # an interior rank's extended strip


def run():
    with torch.no_grad():
        return shard.sharded_filtering_staged(blocks, strips, 0, 1, runners=[shard.CudaStageRunner(b) for b in blocks])


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(a.reps):
    run()
e1.record()
t_enq = (time.perf_counter() - t0) / a.reps * 1e3
torch.cuda.synchronize()
ev = e0.elapsed_time(e1) / a.reps
lib.glrgtv_profile_enable(1)
for _ in range(a.reps):
    run()
torch.cuda.synchronize()
lib.glrgtv_profile_enable(0)
ms = (ctypes.c_float * 16)(); n = (ctypes.c_int * 16)()
lib.glrgtv_profile_read(ms, n, 16)
print(json.dumps({"ranks_emulated": a.ranks, "strip_rows": [int(x.shape[-2]) for x in strips], "event_ms_per_image": round(ev, 3),
                  "host_enqueue_ms_per_image": round(t_enq, 3), "own_kernel_ms_sum": round(sum(ms[i] for i in range(16)) / a.reps, 3),
                  "launches_per_image": int(sum(n[i] for i in range(16)) // a.reps)}))
