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
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
