"""per-phase cycle accounting of the backward stage kernels (debug build with -DGLR_PHASE_TIMING)"""
import ctypes, os, sys, torch
sys.path.insert(0, os.getcwd())
from imagerestoration_development_unrolling_b200 import _lib as L
from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
lib = ctypes.CDLL(os.environ["GLRGTV_LIB"])
dev = torch.device("cuda")
blk = M.LocalLowpassFilteringBlock(48, 1, 8).to(dev)
x = torch.randn(32, 48, 256, 256, device=dev, requires_grad=True); g = torch.randn_like(x)
for _ in range(2):
    blk(x).backward(g)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 64)()
lib.glrgtv_debug_bwd_phases(buf, 1)
blk(x).backward(g); torch.cuda.synchronize()
lib.glrgtv_debug_bwd_phases(buf, 0)
names = ["prologue", "wait_stage", "consume", "phase1", "phase2", "epilogue", "stats_atomics", "-"]
ctas = 32 * 8 * 64
for m, mode in enumerate(["X3", "X2", "X1", "BA"]):
    v = [buf[m * 8 + k] for k in range(8)]
    tot = sum(v)
    print(mode, "cycles/CTA", tot // ctas, {names[k]: f"{100 * v[k] / tot:.1f}%" for k in range(7)})
