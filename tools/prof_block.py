"""One LocalLowpassFilteringBlock forward+backward at a bench shape: the target of ncu captures.
    python tools/prof_block.py [scale 0..3] [batch] [steps]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
s = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
DIMS, NG = [48, 96, 192, 384], [8, 16, 16, 32]
dev = torch.device("cuda")
torch.manual_seed(0)
blk = M.LocalLowpassFilteringBlock(DIMS[s], 1, NG[s]).to(dev)
x = torch.randn(B, DIMS[s], 256 >> s, 256 >> s, device=dev).requires_grad_(True)
g = torch.randn_like(x)
for _ in range(steps):
    out = blk(x)
    torch.autograd.backward([out], [g], inputs=[x] + list(blk.parameters()))
    for p in blk.parameters():
        p.grad = None
    x.grad = None
torch.cuda.synchronize()
print("done")
