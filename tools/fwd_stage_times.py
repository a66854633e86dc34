"""Forward-stage timings of the fused block per scale (CUDA events around each kernel via glrgtv_profile_*).
usage: [GLRGTV_LIB=variant.so] python tools/fwd_stage_times.py [--batch 32] [--bwd]"""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagerestoration_development_unrolling_b200 import _lib as L
from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--bwd", action="store_true")
ap.add_argument("--res", type=int, default=256, help="network-input resolution (scale s runs at res >> s)")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--path", type=int, default=0)
ap.add_argument("--scales", default="0,1,2,3")
ap.add_argument("--tma", type=int, default=0, help="0 auto, 1 cp.async, 2 TMA")
ap.add_argument("--fw2", action="store_true", help="forward stages on the pair walkers (csrc/fw2.cuh)")
a = ap.parse_args()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
lib = L.load()
lib.glrgtv_set_block_path(a.path)
lib.glrgtv_set_stream_loader(a.tma)
lib.glrgtv_set_fwd_kernels(2 if a.fw2 else 1)
SLOTS = ["fwd_weights", "fwd_BA", "fwd_X1", "fwd_X2", "fwd_X3", "bwd_X3", "bwd_X2", "bwd_X1", "bwd_BA", "bwd_weights"]
dev = torch.device("cuda")
out = {}
for s, (d, g) in enumerate(zip([48, 96, 192, 384], [8, 16, 16, 32])):
    if str(s) not in a.scales.split(","):
        continue
    torch.manual_seed(0)
    blk = M.LocalLowpassFilteringBlock(d, 1, g).to(dev)
    x = torch.randn(a.batch, d, a.res >> s, a.res >> s, device=dev, requires_grad=a.bwd)
    go = torch.randn_like(x)
    def run():
        y = blk(x)
        if a.bwd:
            y.backward(go)
            blk.zero_grad(); x.grad = None
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    lib.glrgtv_profile_enable(1)
    for _ in range(a.reps):
        run()
    torch.cuda.synchronize()
    lib.glrgtv_profile_enable(0)
    ms = (ctypes.c_float * 16)(); n = (ctypes.c_int * 16)()
    lib.glrgtv_profile_read(ms, n, 16)
    out[f"scale{s}"] = {SLOTS[i]: round(ms[i] / a.reps, 3) for i in range(len(SLOTS)) if n[i]}
print(json.dumps({"lib": os.environ.get("GLRGTV_LIB", "default"), "batch": a.batch, "res": a.res, "path": a.path, "tma": a.tma, "fw2": a.fw2, **out}))
