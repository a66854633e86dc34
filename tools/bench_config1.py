"""BASELINE config 1 ("GGTV_GGLR_v1.0 denoising, synthetic 128x128 RGB patches sigma=25, batch 4, inference"): latency of
the hot path - the four LocalLowpassFilteringBlock forward passes on the feature maps of a 4x3x128x128 batch
([4,48,128,128], [4,96,64,64], [4,192,32,32], [4,384,16,16]) - on one GPU, eager and replayed from a CUDA graph.
Small shapes live in L2, so this is a latency figure, not a bandwidth one.  (The CPU figure beside it in
profiles/r01_configs.md comes from bench.py's cpu_baseline machinery; tools never import oracle/.)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M  # noqa: E402

DIMS, NG, B, RES = [48, 96, 192, 384], [8, 16, 16, 32], 4, 128
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
blocks = [M.LocalLowpassFilteringBlock(d, 1, g) for d, g in zip(DIMS, NG)]
xs = [torch.randn(B, d, RES >> s, RES >> s) for s, d in enumerate(DIMS)]
gpu = [b.cuda() for b in blocks]
xg = [x.cuda() for x in xs]
with torch.no_grad():
    for _ in range(5):
        outs = [b(x) for b, x in zip(gpu, xg)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        outs = [b(x) for b, x in zip(gpu, xg)]
    e1.record()
    torch.cuda.synchronize()
    ms_gpu = e0.elapsed_time(e1) / n
    # the same forward captured in a CUDA graph (launch-bound at this size: 48 small kernels)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs_g = [b(x) for b, x in zip(gpu, xg)]
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms_graph = e0.elapsed_time(e1) / n
    # the four blocks on four streams (V1X0:1117-1131: they are independent), eager and as one CUDA graph with four branches
    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    ms_streams = timed(lambda: M.run_blocks_on_streams(gpu, xg))
    g4 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g4):
        outs_4 = M.run_blocks_on_streams(gpu, xg)
    ms_graph4 = timed(g4.replay)
    same = all(torch.equal(a, b) for a, b in zip(outs_4, outs_g))
    per_block = [timed(lambda b=b, x=x: b(x)) for b, x in zip(gpu, xg)]
pix = B * RES * RES
print(json.dumps({"config": "config 1: four filter blocks, forward, 4x3x128x128", "gpu_ms": ms_gpu, "gpu_cuda_graph_ms": ms_graph,
                  "gpu_four_streams_ms": ms_streams, "gpu_four_streams_cuda_graph_ms": ms_graph4, "four_streams_bit_equal": same,
                  "per_block_eager_ms": per_block,
                  "gpu_Mpix_per_s": pix / ms_gpu / 1e3, "gpu_graph_Mpix_per_s": pix / ms_graph / 1e3,
                  "finite": bool(all(torch.isfinite(o).all() for o in outs_g))}))
