"""BASELINE config 1 ("GGTV_GGLR_v1.0 denoising, synthetic 128x128 RGB patches sigma=25, batch 4, inference"): latency of
the hot path - the four LocalLowpassFilteringBlock forward passes on the feature maps of a 4x3x128x128 batch
([4,48,128,128], [4,96,64,64], [4,192,32,32], [4,384,16,16]) - on one GPU, next to the oracle port on the host cores.
Small shapes live in L2, so this is a latency figure, not a bandwidth one."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M  # noqa: E402
from oracle import glr_gtv_oracle as O  # noqa: E402

DIMS, NG, B, RES = [48, 96, 192, 384], [8, 16, 16, 32], 4, 128
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
blocks = [M.LocalLowpassFilteringBlock(d, 1, g) for d, g in zip(DIMS, NG)]
states = [{k: v.detach().clone() for k, v in b.state_dict().items()} for b in blocks]
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
    torch.set_num_threads(os.cpu_count() or 1)
    O.lowpass_block_forward(states[3], xs[3])
    t0 = time.perf_counter()
    ref = [O.lowpass_block_forward(sd, x) for sd, x in zip(states, xs)]
    ms_cpu = (time.perf_counter() - t0) * 1e3
err = max(float((o.cpu().double() - r.double()).norm() / r.double().norm()) for o, r in zip(outs_g, ref))
pix = B * RES * RES
print(json.dumps({"config": "config 1: four filter blocks, forward, 4x3x128x128", "gpu_ms": ms_gpu, "gpu_cuda_graph_ms": ms_graph,
                  "gpu_Mpix_per_s": pix / ms_gpu / 1e3, "gpu_graph_Mpix_per_s": pix / ms_graph / 1e3,
                  "cpu_port_ms": ms_cpu, "cpu_cores": os.cpu_count(), "cpu_Mpix_per_s": pix / ms_cpu / 1e3,
                  "max_rel_err_vs_cpu_port": err}))
