"""torch.profiler table of two bench steps (all four blocks, batch 32): per-kernel device times, own kernels and library glue."""
import sys, os, torch
sys.path.insert(0, os.getcwd())
from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
DIMS, NG = [48, 96, 192, 384], [8, 16, 16, 32]
dev = torch.device("cuda")
blocks = [M.LocalLowpassFilteringBlock(d, 1, g).to(dev) for d, g in zip(DIMS, NG)]
params = [p for b in blocks for p in b.parameters()]
xs = [torch.randn(32, d, 256 >> s, 256 >> s, device=dev).requires_grad_(True) for s, d in enumerate(DIMS)]
gs = [torch.randn_like(x) for x in xs]
def step():
    outs = [b(x) for b, x in zip(blocks, xs)]
    torch.autograd.backward(outs, gs, inputs=list(xs) + params)
    for p in params: p.grad = None
    for x in xs: x.grad = None
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
rows = [(e.key, e.self_device_time_total / 2e3, e.count // 2) for e in prof.key_averages() if e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device time per step {tot:.2f} ms")
for k, ms, n in rows:
    print(f"{ms:8.3f} ms  {n:4d}x  {k[:110]}")
